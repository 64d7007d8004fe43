// BOW visual-word assignment + tf-idf scoring, variance-of-Laplacian, particle-filter update.
#pragma once
#include "common.cuh"

namespace mcl {

// ------------------------------------------------------------------------------------------
// BOW vq:  words[i] = argmin_j ||o_i - c_j||^2 (lowest j on ties)   -- scipy.cluster.vq.vq at Matcher.py:250
// fp32 direct-difference form (most accurate fp32 evaluation; differs from scipy's expansion form only on
// near ties).  grid = (word chunks of 64, query row blocks of 128, locations); the winner is merged with a
// 64-bit atomicMin on (distance bits << 32 | word index), which is exactly "smallest distance, lowest index".
// ------------------------------------------------------------------------------------------
constexpr int VQ_WORDS = 64;

__global__ void __launch_bounds__(128) bow_vq_kernel(
    const float4* __restrict__ q_desc /* (n_q,128) f32 */, int n_q, const float4* __restrict__ voc,
    const int64_t* __restrict__ voc_off /* rows, n_loc+1 */, const int* __restrict__ only_if /* optional device flag */,
    int gx /* word chunks */, int gy /* row blocks */, int gz /* locations */,
    unsigned long long* __restrict__ best /* [n_loc][n_q] */) {
    // as the fallback of the tensor-core path the kernel is launched with a small grid that walks the (gx, gy, gz) work
    // space: when the device flag says "not needed" that is a few hundred empty blocks instead of a few hundred thousand
    if (only_if && *only_if == 0) return;
    __shared__ float4 sm[VQ_WORDS][32];
    const int64_t n_work = (int64_t)gx * gy * gz;
    for (int64_t work = blockIdx.x; work < n_work; work += gridDim.x) {
        const int bx = (int)(work % gx), by = (int)((work / gx) % gy), loc = (int)(work / ((int64_t)gx * gy));
        const int64_t v0 = voc_off[loc];
        const int n_words = (int)(voc_off[loc + 1] - v0);
        const int w0 = bx * VQ_WORDS;
        if (w0 >= n_words) continue;   // block-uniform
        const int nw = min(VQ_WORDS, n_words - w0);
        __syncthreads();               // the previous work item's readers are done with sm
        for (int i = threadIdx.x; i < nw * 32; i += blockDim.x) sm[i >> 5][i & 31] = voc[(v0 + w0) * 32 + i];
        __syncthreads();
        const int row = by * blockDim.x + threadIdx.x;
        if (row < n_q) {
            float4 a[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) a[c] = q_desc[(int64_t)row * 32 + c];
            float bd = INFINITY;
            int bj = 0;
            for (int j = 0; j < nw; ++j) {
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const float4 b = sm[j][c];
                    const float dx = a[c].x - b.x, dy = a[c].y - b.y, dz = a[c].z - b.z, dw = a[c].w - b.w;
                    s0 = fmaf(dx, dx, s0); s1 = fmaf(dy, dy, s1); s2 = fmaf(dz, dz, s2); s3 = fmaf(dw, dw, s3);
                }
                const float d = (s0 + s1) + (s2 + s3);
                if (d < bd) { bd = d; bj = w0 + j; }
            }
            const unsigned long long key = ((unsigned long long)__float_as_uint(bd) << 32) | (unsigned)bj;
            atomicMin(&best[(int64_t)loc * n_q + row], key);
        }
    }
}

__device__ __forceinline__ void bow_gather_scores(const int* __restrict__ nz_word, const float* __restrict__ nz_val, int n_nz,
                                                  const float* __restrict__ im_features_t, const int64_t* __restrict__ img_off,
                                                  int loc, int W, int frame, int total_images, float* __restrict__ out_scores);

// tf-idf scoring (Matcher.py:249-258): histogram of words, *idf, L2 normalise, dot with every image row.
// The query vector has at most (rows of the frame) non-zeros out of W = 10000 words, so after normalisation it is
// compacted, in word order (deterministic summation order), into a (word, value) list and every image row is only
// gathered at those words: ~300 x 100 gathers per location instead of a dense 4 MB read.
// grid = (n_locations, n_frames), block = 256; dynamic smem = W floats + list_cap x (int + float).
__global__ void __launch_bounds__(256) bow_score_kernel(
    const unsigned long long* __restrict__ best /* [n_loc][n_q] */, int n_q, const int64_t* __restrict__ q_off /* frames */,
    const float* __restrict__ idf /* [n_loc][W] */, const float* __restrict__ im_features_t /* per location (W, N_l) */,
    const int64_t* __restrict__ img_off /* n_loc+1 */, int W, int total_images, int list_cap,
    int32_t* __restrict__ out_words, float* __restrict__ out_scores /* [n_frames][total_images] */) {
    extern __shared__ float hist[];
    int* nz_word = reinterpret_cast<int*>(hist + W);
    float* nz_val = reinterpret_cast<float*>(nz_word + list_cap);
    __shared__ float red[8];
    __shared__ float inv_norm_s;
    __shared__ int warp_cnt[8];
    __shared__ int nz_total;
    const int loc = blockIdx.x, frame = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int w = threadIdx.x; w < W; w += blockDim.x) hist[w] = 0.f;
    if (threadIdx.x == 0) nz_total = 0;
    __syncthreads();
    const int64_t q0 = q_off[frame], q1 = q_off[frame + 1];
    for (int64_t i = q0 + threadIdx.x; i < q1; i += blockDim.x) {
        const int word = (int)(best[(int64_t)loc * n_q + i] & 0xffffffffu);
        if (out_words) out_words[(int64_t)loc * n_q + i] = word;
        atomicAdd(&hist[word], 1.0f);  // integer-valued floats: order independent
    }
    __syncthreads();
    float ss = 0.f;
    for (int w = threadIdx.x; w < W; w += blockDim.x) {
        const float v = hist[w] * idf[(int64_t)loc * W + w];
        hist[w] = v;
        ss += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        const float nrm = sqrtf(t);
        inv_norm_s = nrm > 0.f ? nrm : 1.0f;  // sklearn normalize: zero rows are left unchanged
    }
    __syncthreads();
    const float nrm = inv_norm_s;
    // ordered compaction of the non-zeros
    for (int base = 0; base < W; base += blockDim.x) {
        const int w = base + threadIdx.x;
        const float v = w < W ? hist[w] : 0.f;
        const bool nz = v != 0.f;
        const unsigned bal = __ballot_sync(0xffffffffu, nz);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int off = nz_total;
        for (int k = 0; k < warp; ++k) off += warp_cnt[k];
        if (nz) {
            const int pos = off + __popc(bal & ((1u << lane) - 1u));
            nz_word[pos] = w;
            nz_val[pos] = v / nrm;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int k = 0; k < 8; ++k) t += warp_cnt[k];
            nz_total += t;
        }
        __syncthreads();
    }
    const int n_nz = nz_total;
    bow_gather_scores(nz_word, nz_val, n_nz, im_features_t, img_off, loc, W, frame, total_images, out_scores);
}

// Scores of one (location, frame) from its compacted query vector: im_features_t holds every location's tf-idf index
// TRANSPOSED, (W, N_l) at offset img_off[l] * W, so the N_l image values of a word are contiguous (a (N_l, W) layout costs a
// 32-byte sector per 4 useful bytes: 8x the traffic).  The list is dealt to the 8 warps (entry k to warp k mod 8); a lane owns
// images lane, lane + 32, lane + 64, lane + 96 of a 128-image chunk, so a word costs one coalesced row per warp and a warp has
// 4 x its unroll depth of independent loads in flight (thread = image with the whole list in sequence was one dependent chain
// of ~300 loads: 0.145 ms for 50 frames x 32 locations).  The 8 partial sums of an image are added in warp order: the
// summation order is fixed, so the scores are reproducible run to run (and equal between the host and the resident path).
// block = 256.
__device__ __forceinline__ void bow_gather_scores(const int* __restrict__ nz_word, const float* __restrict__ nz_val, int n_nz,
                                                  const float* __restrict__ im_features_t, const int64_t* __restrict__ img_off,
                                                  int loc, int W, int frame, int total_images, float* __restrict__ out_scores) {
    __shared__ float partial[8][128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t i0 = img_off[loc];
    const int n_img = (int)(img_off[loc + 1] - i0);
    const float* base = im_features_t + i0 * W;
    for (int chunk = 0; chunk < n_img; chunk += 128) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        bool on[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) on[j] = chunk + lane + 32 * j < n_img;
#pragma unroll 4
        for (int k = warp; k < n_nz; k += 8) {
            const float val = nz_val[k];
            const float* row = base + (int64_t)nz_word[k] * n_img + chunk + lane;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (on[j]) acc[j] = fmaf(val, row[32 * j], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) partial[warp][lane + 32 * j] = acc[j];
        __syncthreads();
        if (threadIdx.x < 128 && chunk + (int)threadIdx.x < n_img) {
            float t = partial[0][threadIdx.x];
#pragma unroll
            for (int g = 1; g < 8; ++g) t += partial[g][threadIdx.x];
            out_scores[(int64_t)frame * total_images + i0 + chunk + threadIdx.x] = t;
        }
        __syncthreads();
    }
}

// The same scoring for frames of up to BOW_SPARSE_ROWS descriptors without the dense W-long passes (zeroing, idf multiply, norm
// and compaction over 10 000 words for ~300 occupied ones): every row enters its word into a small open-addressing hash table in
// shared memory (atomicCAS on the key, then a count and the smallest row index per word), and the first row of a word carries
// count * idf; the list is compacted in row order.  (Comparing every row with all rows, the first version, was 60 % of the
// kernel's instructions.)
// grid = (n_locations, n_frames), block = 256.
constexpr int BOW_SPARSE_ROWS = 512;
constexpr int BOW_HASH_BITS = 10;
constexpr int BOW_HASH = 1 << BOW_HASH_BITS;   // 2 x BOW_SPARSE_ROWS slots

__global__ void __launch_bounds__(256) bow_score_sparse_kernel(
    const unsigned long long* __restrict__ best /* [n_loc][n_q] */, int n_q, const int64_t* __restrict__ q_off /* frames */,
    const float* __restrict__ idf /* [n_loc][W] */, const float* __restrict__ im_features_t, const int64_t* __restrict__ img_off,
    int W, int total_images, int32_t* __restrict__ out_words, float* __restrict__ out_scores) {
    __shared__ int wd[BOW_SPARSE_ROWS];
    __shared__ int h_key[BOW_HASH], h_cnt[BOW_HASH], h_first[BOW_HASH];
    __shared__ int nz_word[BOW_SPARSE_ROWS];
    __shared__ float nz_val[BOW_SPARSE_ROWS];
    __shared__ float red[8];
    __shared__ int warp_cnt[8];
    __shared__ float nrm_s;
    __shared__ int nz_total;
    const int loc = blockIdx.x, frame = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t q0 = q_off[frame];
    const int n = (int)(q_off[frame + 1] - q0);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int word = (int)(best[(int64_t)loc * n_q + q0 + i] & 0xffffffffu);
        wd[i] = word;
        if (out_words) out_words[(int64_t)loc * n_q + q0 + i] = word;
    }
    for (int i = threadIdx.x; i < BOW_HASH; i += blockDim.x) {
        h_key[i] = -1;
        h_cnt[i] = 0;
        h_first[i] = INT_MAX;
    }
    if (threadIdx.x == 0) nz_total = 0;
    __syncthreads();
    int slot[BOW_SPARSE_ROWS / 256];
#pragma unroll
    for (int r = 0; r < BOW_SPARSE_ROWS / 256; ++r) {
        const int i = r * 256 + threadIdx.x;
        slot[r] = -1;
        if (i < n) {
            const int w = wd[i];
            int sl = (int)(((unsigned)w * 2654435761u) >> (32 - BOW_HASH_BITS));
            while (true) {   // n <= BOW_SPARSE_ROWS = half the table: a free or matching slot always exists
                const int prev = atomicCAS(&h_key[sl], -1, w);
                if (prev == -1 || prev == w) break;
                sl = (sl + 1) & (BOW_HASH - 1);
            }
            atomicAdd(&h_cnt[sl], 1);
            atomicMin(&h_first[sl], i);
            slot[r] = sl;
        }
    }
    __syncthreads();
    float v[BOW_SPARSE_ROWS / 256];
    float ss = 0.f;
#pragma unroll
    for (int r = 0; r < BOW_SPARSE_ROWS / 256; ++r) {
        const int i = r * 256 + threadIdx.x;
        v[r] = 0.f;
        if (slot[r] >= 0 && h_first[slot[r]] == i)   // histogram count (exact) * idf, float32 like the reference
            v[r] = (float)h_cnt[slot[r]] * idf[(int64_t)loc * W + wd[i]];
        ss += v[r] * v[r];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        const float nrm = sqrtf(t);
        nrm_s = nrm > 0.f ? nrm : 1.0f;   // sklearn normalize: zero rows are left unchanged
    }
    __syncthreads();
    const float nrm = nrm_s;
#pragma unroll
    for (int r = 0; r < BOW_SPARSE_ROWS / 256; ++r) {   // ordered compaction of the non-zeros (row order)
        const int i = r * 256 + threadIdx.x;
        const bool nz = v[r] != 0.f;
        const unsigned bal = __ballot_sync(0xffffffffu, nz);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int off = nz_total;
        for (int k = 0; k < warp; ++k) off += warp_cnt[k];
        if (nz) {
            const int pos = off + __popc(bal & ((1u << lane) - 1u));
            nz_word[pos] = wd[i];
            nz_val[pos] = v[r] / nrm;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int k = 0; k < 8; ++k) t += warp_cnt[k];
            nz_total += t;
        }
        __syncthreads();
    }
    bow_gather_scores(nz_word, nz_val, nz_total, im_features_t, img_off, loc, W, frame, total_images, out_scores);
}

// (N_l, W) -> (W, N_l) per location, once per index.  grid = (ceil(W / 32), n_locations), block = (32, 8).
__global__ void bow_transpose_imf_kernel(const float* __restrict__ imf, const int64_t* __restrict__ img_off, int W,
                                         float* __restrict__ imf_t) {
    __shared__ float tile[32][33];
    const int loc = blockIdx.y;
    const int64_t i0 = img_off[loc];
    const int n_img = (int)(img_off[loc + 1] - i0);
    const int w0 = blockIdx.x * 32;
    for (int b = 0; b < n_img; b += 32) {
        for (int r = threadIdx.y; r < 32; r += 8) {
            const int im = b + r, w = w0 + threadIdx.x;
            tile[r][threadIdx.x] = (im < n_img && w < W) ? imf[(i0 + im) * W + w] : 0.f;
        }
        __syncthreads();
        for (int r = threadIdx.y; r < 32; r += 8) {
            const int w = w0 + r, im = b + threadIdx.x;
            if (w < W && im < n_img) imf_t[i0 * W + (int64_t)w * n_img + im] = tile[threadIdx.x][r];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// variance of the Laplacian (analyze.py:441-445): 4-neighbour kernel, BORDER_REFLECT_101, population var.
// One byte of HBM traffic per pixel, so the instruction count per pixel decides whether the kernel can follow HBM:
// at 6.5 TB/s the SMs have ~5.6 issue slots per pixel.  The arithmetic is therefore done on PAIRS of pixels held as two
// 16-bit lanes of one register (every intermediate is non-negative and < 2^16, so plain 32-bit adds never carry across
// the lanes):
//     P[j]  = (px[2j], px[2j+1])                      one PRMT per pair, per row, reused as up / mid / down row
//     Q[j]  = (px[2j-1], px[2j])                      one PRMT per pair of the mid row (odd-aligned pairs)
//     Lb[j] = upP + dnP + Q[j] + Q[j+1] + 1020 - 4*midP     two IADD3 + one IMAD per pair;  Lb = L + 1020 in [0, 2040]
// then one lane extraction and one IMAD per pixel for sum Lb^2.  sum L needs no per-pixel work at all: the stencil's
// coefficients add up to zero, so the sum over the image telescopes to the two outermost rows and columns
// (laplacian_finish_kernel), which also removes the bias exactly: sum Lb = sum L + 1020 N,
// sum L^2 = sum Lb^2 - 2040 sum Lb + 1020^2 N.  ~4.6 instructions per pixel instead of ~12 for the scalar form.
// A thread owns a 16-pixel-wide column block (one aligned 128-bit load per row; rows are stored with a 16-byte pitch)
// over LAP_RY rows; the two halo pixels come from the neighbouring lanes by shuffle (byte loads only at warp edges).
// Threads are numbered (row strip, column block) within an image so that no lane idles on widths like 800 = 50 * 16.
// grid = (ceil(strips * col_blocks / 256), n_images), block = 256.
// ------------------------------------------------------------------------------------------
constexpr int LAP_RY = 16;   // rows per thread
constexpr int LAP_PF = 4;    // rows loaded per batch (loads in flight per thread)
constexpr uint32_t LAP_BIAS = 1020u | (1020u << 16);

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

constexpr int LAP_PX = 16;   // pixels per thread and row (one aligned 128-bit load; 8 = 64-bit loads measured 40 % slower)
constexpr int LAP_NW = LAP_PX / 4, LAP_NP = LAP_PX / 2;

struct LapRow {
    uint32_t p[LAP_NP];   // p[j] = px[2j] | px[2j+1] << 16
    uint32_t lh, rh;      // pixel left of px[0] / right of px[PX-1] (border-reflected), as plain integers
};
struct LapWords {
    uint32_t w[LAP_NW];
};

__device__ __forceinline__ void lap_unpack(const LapWords& v, LapRow& r) {
#pragma unroll
    for (int i = 0; i < LAP_NW; ++i) {
        r.p[2 * i] = __byte_perm(v.w[i], 0u, 0x4140);
        r.p[2 * i + 1] = __byte_perm(v.w[i], 0u, 0x4342);
    }
}

// lane extraction on the FMA pipe (IMAD.HI) instead of a shift on the ALU pipe, which carries most of this kernel's work
__device__ __forceinline__ uint32_t lap_hi16(uint32_t x) {
    uint32_t r;
    asm("mul.hi.u32 %0, %1, 65536;" : "=r"(r) : "r"(x));
    return r;
}
// MASKED: the image width is not a multiple of LAP_PX, i.e. the last column block has pixels outside the image
template <bool MASKED>
__device__ __forceinline__ void lap_body(const uint8_t* __restrict__ img, int H, int W, int pitch, uint32_t& s2) {
    constexpr int PX = LAP_PX, NP = LAP_NP, NW = LAP_NW;
    const int CB = (W + PX - 1) / PX, NS = (H + LAP_RY - 1) / LAP_RY;
    const int t = blockIdx.x * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool active = t < CB * NS;
    const int strip = active ? t / CB : NS - 1, cb = active ? t % CB : CB - 1;
    const int x0 = cb * PX, ys = strip * LAP_RY;
    const int k = MASKED ? min(PX, W - x0) : PX;    // valid pixels of this column block
    const bool from_left = cb > 0 && lane > 0;      // lane-1 holds the column block to the left, same rows
    const bool from_right = cb < CB - 1 && lane < 31;
    const bool load_left = cb > 0 && lane == 0, load_right = cb < CB - 1 && lane == 31;   // halo not in a neighbouring lane

    uint32_t m[NP];   // MASKED: which lanes of each pair are pixels of the image
    if (MASKED) {
#pragma unroll
        for (int j = 0; j < NP; ++j) m[j] = (2 * j + 1 < k) ? 0xffffffffu : (2 * j < k ? 0x0000ffffu : 0u);
    }

    // rows are clamped into the image: rows past the strip's end are loaded (harmlessly) but never accumulated.
    // issue(): the loads only, so that a batch of rows is in flight before anything waits on one of them;
    // halo(): the neighbour pixels, from the adjacent lanes' words where possible
    auto issue = [&](int y, LapWords& v, uint32_t& bl, uint32_t& br) {
        const uint8_t* row = img + (int64_t)min(max(y, 0), H - 1) * pitch;
        if (NW == 4) {
            const uint4 x = *reinterpret_cast<const uint4*>(row + x0);
            v.w[0] = x.x; v.w[1] = x.y; v.w[NW - 2] = x.z; v.w[NW - 1] = x.w;
        } else {
            const uint2 x = *reinterpret_cast<const uint2*>(row + x0);
            v.w[0] = x.x; v.w[NW - 1] = x.y;
        }
        bl = 0;
        br = 0;
        if (load_left) bl = row[x0 - 1];
        if (load_right) br = row[x0 + PX];
    };
    auto halo = [&](const LapWords& v, uint32_t bl, uint32_t br, uint32_t& lh, uint32_t& rh) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, v.w[NW - 1], 1), dn = __shfl_down_sync(0xffffffffu, v.w[0], 1);
        lh = from_left ? up >> 24 : bl;
        rh = from_right ? dn & 0xffu : br;
        if (cb == 0) lh = W > 1 ? (v.w[0] >> 8) & 0xffu : v.w[0] & 0xffu;   // REFLECT_101: pixel -1 mirrors pixel 1
        if (cb == CB - 1) rh = (v.w[NW - 1] >> 16) & 0xffu;                 // k == PX: pixel W mirrors pixel W-2 (else patched below)
    };
    auto finish_row = [&](const LapWords& v, uint32_t lh, uint32_t rh, LapRow& r) {
        lap_unpack(v, r);
        r.lh = lh;
        r.rh = rh;
        if (MASKED && k < PX) {  // pixel k (the first one outside the image) mirrors pixel k-2
            const int kj = k >> 1;
            const uint32_t lm = (k & 1) ? 0xffff0000u : 0x0000ffffu;
#pragma unroll
            for (int j = 1; j < NP; ++j)
                if (j == kj) r.p[j] = (r.p[j] & ~lm) | (r.p[j - 1] & lm);
            if (k == 1) r.p[0] = (r.p[0] & 0xffffu) | ((W > 1 ? lh : (r.p[0] & 0xffffu)) << 16);
        }
    };

    LapRow up, mid, dn;
    {
        LapWords v0, v1;
        uint32_t bl0, br0, bl1, br1, l0, r0, l1, r1;
        issue(reflect101(ys - 1, H), v0, bl0, br0);
        issue(ys, v1, bl1, br1);
        halo(v0, bl0, br0, l0, r0);
        halo(v1, bl1, br1, l1, r1);
        finish_row(v0, l0, r0, up);
        finish_row(v1, l1, r1, mid);
    }
#pragma unroll 1
    for (int g = 0; g < LAP_RY / LAP_PF; ++g) {
        LapWords v[LAP_PF];
        uint32_t bl[LAP_PF], br[LAP_PF], lh[LAP_PF], rh[LAP_PF];
#pragma unroll
        for (int r = 0; r < LAP_PF; ++r) {
            const int y = ys + g * LAP_PF + r;   // the row whose Laplacian this step computes; loads row y+1
            issue(y + 1 < H ? y + 1 : reflect101(y + 1, H), v[r], bl[r], br[r]);
        }
#pragma unroll
        for (int r = 0; r < LAP_PF; ++r) halo(v[r], bl[r], br[r], lh[r], rh[r]);
#pragma unroll
        for (int r = 0; r < LAP_PF; ++r) {
            const int y = ys + g * LAP_PF + r;
            finish_row(v[r], lh[r], rh[r], dn);
            uint32_t q[NP + 1];
            q[0] = __byte_perm(mid.lh, mid.p[0], 0x5410);
#pragma unroll
            for (int j = 1; j < NP; ++j) q[j] = __byte_perm(mid.p[j - 1], mid.p[j], 0x5432);
            q[NP] = __byte_perm(mid.p[NP - 1], mid.rh, 0x5432);
            uint32_t rs2 = 0;
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const uint32_t qq = q[j] + q[j + 1] + LAP_BIAS;
                uint32_t lb = up.p[j] + dn.p[j] + qq;
                lb -= 4u * mid.p[j];
                if (MASKED) lb &= m[j];
                const uint32_t lo = lb & 0xffffu, hi = lap_hi16(lb);
                rs2 += lo * lo;
                rs2 += hi * hi;
            }
            if (active && y < H) s2 += rs2;
            up = mid;
            mid = dn;
        }
    }
}

__global__ void __launch_bounds__(256) laplacian_sums_kernel(const uint8_t* __restrict__ pixels,
                                                             const int64_t* __restrict__ img_off,
                                                             const int* __restrict__ pitches,
                                                             const int* __restrict__ hs, const int* __restrict__ ws,
                                                             unsigned long long* __restrict__ sums /* [n][2]: -, sum Lb^2 */) {
    const int im = blockIdx.y;
    const int H = hs[im], W = ws[im];
    uint32_t s2 = 0;   // per thread: LAP_PX * LAP_RY = 256 pixels * 2040^2 (1.07e9) fits 32 bits
    if (W % LAP_PX) lap_body<true>(pixels + img_off[im], H, W, pitches[im], s2);
    else lap_body<false>(pixels + img_off[im], H, W, pitches[im], s2);
    unsigned long long t2 = s2;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t2 += __shfl_xor_sync(0xffffffffu, t2, o);
    if ((threadIdx.x & 31) == 0 && t2 != 0) atomicAdd(&sums[2 * im + 1], t2);
}

// One warp per image.  sum L: with E the REFLECT_101 extension of the image I,
//     sum_{y,x} (E(y-1,x) + E(y+1,x) - 2 I(y,x)) = sum_x  I(r(1),x) - I(0,x) + I(r(H-2),x) - I(H-1,x)      (telescoping in y)
// and likewise in x, r() / c() being the reflected row / column index -- O(H + W) pixels.  Then the +1020 bias of the
// main kernel is removed exactly (64-bit integers): sums[2i] = sum L, sums[2i+1] = sum L^2, and
// var = (sum L^2 - (sum L)^2 / N) / N in float64 (the formula SURVEY 8(a) found to match ndarray.var() to 1 ulp).
__global__ void __launch_bounds__(128) laplacian_finish_kernel(const uint8_t* __restrict__ pixels, const int64_t* __restrict__ img_off,
                                                               const int* __restrict__ pitches, const int* __restrict__ hs,
                                                               const int* __restrict__ ws, int n, long long* __restrict__ sums,
                                                               double* __restrict__ out) {
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    const int H = hs[i], W = ws[i], pitch = pitches[i];
    const uint8_t* img = pixels + img_off[i];
    const uint8_t *r0 = img, *r1 = img + (int64_t)reflect101(1, H) * pitch, *r2 = img + (int64_t)reflect101(H - 2, H) * pitch,
                  *r3 = img + (int64_t)(H - 1) * pitch;
    const int c1 = reflect101(1, W), c2 = reflect101(W - 2, W), c3 = W - 1;
    int acc = 0;   // |acc| <= 2 * 255 * (H + W) / 32 per lane
    for (int x = lane; x < W; x += 32) acc += (int)r1[x] - (int)r0[x] + (int)r2[x] - (int)r3[x];
    for (int y = lane; y < H; y += 32) {
        const uint8_t* row = img + (int64_t)y * pitch;
        acc += (int)row[c1] - (int)row[0] + (int)row[c2] - (int)row[c3];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane != 0) return;
    const long long cnt = (long long)H * (long long)W;
    const long long l1 = acc;
    const long long b1 = l1 + 1020LL * cnt, b2 = sums[2 * i + 1];
    const long long l2 = b2 - 2040LL * b1 + 1020LL * 1020LL * cnt;
    sums[2 * i] = l1;
    sums[2 * i + 1] = l2;
    const double N = (double)cnt, s1 = (double)l1, s2 = (double)l2;
    out[i] = __ddiv_rn(__dsub_rn(s2, __ddiv_rn(__dmul_rn(s1, s1), N)), N);
}

// ------------------------------------------------------------------------------------------
// particle-filter update for one frame (analyze.py:119-134): accountCommand -> prevWeight -> probUpdate ->
// best circle / best angle.  State layout: L rows of (1 + I) doubles = [total, p_0 .. p_{I-1}].
// Arithmetic is float64 with explicit mul/add (no FMA contraction) in the reference's operation order, so the
// result is bit-identical to the Python lists.  One CTA; thread per (location, slot).
// `prev` is updated in place the way the reference's shallow copy mutates previousProbs (analyze.py:317-335).
// fwd_factor: host-computed 0.05*|sin(k*15*180/pi)| table (libm sin, like math.sin), I entries.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int lex_compare(const double* a, const double* b, int n) {
    for (int i = 0; i < n; ++i) {
        if (a[i] < b[i]) return -1;
        if (a[i] > b[i]) return 1;
    }
    return 0;
}
__device__ int lex_argmax_rows(const double* m, int L, int stride) {  // first maximal row, Python max()/index()
    int best = 0;
    for (int l = 1; l < L; ++l)
        if (lex_compare(m + (size_t)l * stride, m + (size_t)best * stride, stride) > 0) best = l;
    return best;
}
__device__ int first_argmax(const double* v, int n) {
    int best = 0;
    for (int i = 1; i < n; ++i)
        if (v[i] > v[best]) best = i;
    return best;
}

// Typed scalars.  The reference's lists hold Python floats/ints, and -- for the colour method -- np.float32 values
// (cv2 histograms -> float32 chi-squared distances) and np.float64 values (the blur weight derives from ndarray.var()).
// Python's operators then follow numpy's scalar promotion: a Python float is a weak scalar (float32 op python -> float32,
// the Python value rounded to float32 first), np.float64 is strong (anything op float64 -> float64).  Rows carry a tag and
// every multiply / add below applies that rule, so the float32 roundings of the reference are reproduced exactly.  Two tags per row: [total, probabilities].
enum { T_PY = 0, T_F64 = 1, T_F32 = 2 };
struct TV { double v; int t; };
__device__ __forceinline__ int t_promote(int a, int b) { return a == b ? a : ((a == T_F64 || b == T_F64) ? T_F64 : T_F32); }
__device__ __forceinline__ TV t_mul(TV a, TV b) {
    TV r; r.t = t_promote(a.t, b.t);
    r.v = (r.t == T_F32) ? (double)__fmul_rn((float)a.v, (float)b.v) : __dmul_rn(a.v, b.v);
    return r;
}
__device__ __forceinline__ TV t_add(TV a, TV b) {
    TV r; r.t = t_promote(a.t, b.t);
    r.v = (r.t == T_F32) ? (double)__fadd_rn((float)a.v, (float)b.v) : __dadd_rn(a.v, b.v);
    return r;
}

__global__ void filter_step_kernel(int L, int I, int step_deg, int stages, int cmd, double blur, int blur_t, double* __restrict__ prev,
                                   const signed char* __restrict__ prev_t, const double* __restrict__ cur,
                                   const signed char* __restrict__ cur_t, const double* __restrict__ fwd_factor,
                                   double* __restrict__ out, signed char* __restrict__ out_t, int* __restrict__ best /* 2 */,
                                   double* __restrict__ scratch /* L*(1+I) */) {
    const int S = 1 + I;
    const int n = L * S;
    // ---- accountCommand
    if (stages & 1) {
        if (cmd == 'l' || cmd == 'r') {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int l = i / S, k = i % S;
                if (k == 0) scratch[i] = prev[i];
                else {
                    const int a = k - 1;
                    const int srcA = (cmd == 'l') ? (a + 1) % I : (a + I - 1) % I;
                    scratch[i] = prev[l * S + 1 + srcA];
                }
            }
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += blockDim.x) prev[i] = scratch[i];
            __syncthreads();
        } else if (cmd == 'f') {
            if (threadIdx.x == 0) {
                const int bc = lex_argmax_rows(prev, L, S);
                const int ba = first_argmax(prev + (size_t)bc * S + 1, I);
                const TV onep = {__dadd_rn(1.0, fwd_factor[ba]), T_PY};
                int row = -1;
                if (bc < L - 1 && ba * step_deg < 180 && ba > 0) row = bc + 1;              // analyze.py:332-335 (15 = the angle step)
                else if (bc > 0 && ba * step_deg > 180 && ba * step_deg < 360) row = bc - 1;
                if (row >= 0) {
                    const TV tot = {prev[(size_t)row * S], prev_t ? prev_t[2 * row] : T_PY};
                    prev[(size_t)row * S] = t_mul(tot, onep).v;   // `*=` with a Python float keeps the row's type
                }
            }
            __syncthreads();
        }
    }
    // ---- prevWeight then probUpdate
    const TV cw1 = {0.7, T_PY}, pw1 = {1 - 0.7, T_PY};
    const bool sharp = blur > 200;
    const int wt = sharp ? T_PY : blur_t;   // 0.85 is a literal; blur/200*0.85 has the blur's type
    const TV cw2 = {sharp ? 0.85 : __dmul_rn(__ddiv_rn(blur, 200.0), 0.85), wt};
    const TV pw2 = {__dsub_rn(1.0, cw2.v), wt};
    if (stages & 6) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int l = i / S, k = i % S, slot = 2 * l + (k != 0);   // tags: [total, probabilities] per row
            const TV p = {prev[i], prev_t ? prev_t[slot] : T_PY};
            TV x = {cur[i], cur_t ? cur_t[slot] : T_PY};
            if (stages & 2) x = t_add(t_mul(cw1, x), t_mul(pw1, p));
            if (stages & 4) x = t_add(t_mul(cw2, x), t_mul(pw2, p));
            out[i] = x.v;
            if (out_t && k <= 1) out_t[slot] = (signed char)x.t;
        }
        __syncthreads();
    }
    if ((stages & 8) && threadIdx.x == 0) {
        const double* src = (stages & 6) ? out : prev;
        const int bc = lex_argmax_rows(src, L, S);
        best[0] = bc;
        best[1] = first_argmax(src + (size_t)bc * S + 1, I);
    }
}

}  // namespace mcl
