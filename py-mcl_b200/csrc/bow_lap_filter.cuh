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
    const int64_t* __restrict__ voc_off /* rows, n_loc+1 */, unsigned long long* __restrict__ best /* [n_loc][n_q] */) {
    const int loc = blockIdx.z;
    const int64_t v0 = voc_off[loc];
    const int n_words = (int)(voc_off[loc + 1] - v0);
    const int w0 = blockIdx.x * VQ_WORDS;
    if (w0 >= n_words) return;
    const int nw = min(VQ_WORDS, n_words - w0);
    __shared__ float4 sm[VQ_WORDS][32];
    for (int i = threadIdx.x; i < nw * 32; i += blockDim.x) sm[i >> 5][i & 31] = voc[(v0 + w0) * 32 + i];
    __syncthreads();
    const int row = blockIdx.y * blockDim.x + threadIdx.x;
    if (row >= n_q) return;
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

// tf-idf scoring (Matcher.py:249-258): histogram of words, *idf, L2 normalise, dot with every image row.
// grid = (n_locations, n_frames), block = 256; dynamic smem = W floats.
__global__ void __launch_bounds__(256) bow_score_kernel(
    const unsigned long long* __restrict__ best /* [n_loc][n_q] */, int n_q, const int64_t* __restrict__ q_off /* frames */,
    const float* __restrict__ idf /* [n_loc][W] */, const float* __restrict__ im_features /* rows of W */,
    const int64_t* __restrict__ img_off /* n_loc+1 */, int W, int total_images, int32_t* __restrict__ out_words,
    float* __restrict__ out_scores /* [n_frames][total_images] */) {
    extern __shared__ float hist[];
    __shared__ float red[8];
    __shared__ float inv_norm_s;
    const int loc = blockIdx.x, frame = blockIdx.y;
    for (int w = threadIdx.x; w < W; w += blockDim.x) hist[w] = 0.f;
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
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        const float nrm = sqrtf(t);
        inv_norm_s = nrm > 0.f ? nrm : 1.0f;  // sklearn normalize: zero rows are left unchanged
    }
    __syncthreads();
    const float nrm = inv_norm_s;
    for (int w = threadIdx.x; w < W; w += blockDim.x) hist[w] = hist[w] / nrm;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t i0 = img_off[loc], i1 = img_off[loc + 1];
    for (int64_t im = i0 + warp; im < i1; im += 8) {
        const float* row = im_features + im * W;
        float acc = 0.f;
        for (int w = lane; w < W; w += 32) acc = fmaf(hist[w], row[w], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) out_scores[(int64_t)frame * total_images + im] = acc;
    }
}

// ------------------------------------------------------------------------------------------
// variance of the Laplacian (analyze.py:441-445): 4-neighbour kernel, BORDER_REFLECT_101, population var.
// Shared-memory staged 32x128 tiles with a 1-pixel halo; exact integer sums (sum L in s64, sum L^2 in u64).
// grid = (tiles_x, tiles_y, n_images).
// ------------------------------------------------------------------------------------------
constexpr int LAP_TW = 128, LAP_TH = 32;

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

__global__ void __launch_bounds__(256) laplacian_sums_kernel(const uint8_t* __restrict__ pixels,
                                                             const int64_t* __restrict__ img_off,
                                                             const int* __restrict__ hs, const int* __restrict__ ws,
                                                             long long* __restrict__ sums /* [n][2] */) {
    const int im = blockIdx.z;
    const int H = hs[im], W = ws[im];
    const int x0 = blockIdx.x * LAP_TW, y0 = blockIdx.y * LAP_TH;
    if (x0 >= W || y0 >= H) return;
    const uint8_t* src = pixels + img_off[im];
    __shared__ uint8_t tile[LAP_TH + 2][LAP_TW + 2 + 2];
    for (int i = threadIdx.x; i < (LAP_TH + 2) * (LAP_TW + 2); i += blockDim.x) {
        const int ty = i / (LAP_TW + 2), tx = i % (LAP_TW + 2);
        const int y = reflect101(y0 + ty - 1, H), x = reflect101(x0 + tx - 1, W);
        tile[ty][tx] = src[(int64_t)y * W + x];
    }
    __syncthreads();
    long long s1 = 0;
    unsigned long long s2 = 0;
    for (int i = threadIdx.x; i < LAP_TH * LAP_TW; i += blockDim.x) {
        const int ty = i / LAP_TW, tx = i % LAP_TW;
        if (y0 + ty < H && x0 + tx < W) {
            const int l = (int)tile[ty][tx + 1] + (int)tile[ty + 2][tx + 1] + (int)tile[ty + 1][tx] +
                          (int)tile[ty + 1][tx + 2] - 4 * (int)tile[ty + 1][tx + 1];
            s1 += l;
            s2 += (unsigned long long)(l * l);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(reinterpret_cast<unsigned long long*>(&sums[2 * im]), (unsigned long long)s1);
        atomicAdd(reinterpret_cast<unsigned long long*>(&sums[2 * im + 1]), s2);
    }
}

__global__ void laplacian_var_kernel(const long long* __restrict__ sums, const int* __restrict__ hs,
                                     const int* __restrict__ ws, int n, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double N = (double)hs[i] * (double)ws[i];
    const double s1 = (double)sums[2 * i], s2 = (double)sums[2 * i + 1];
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

__global__ void filter_step_kernel(int L, int I, int stages, int cmd, double blur, double* __restrict__ prev,
                                   const double* __restrict__ cur, const double* __restrict__ fwd_factor,
                                   double* __restrict__ out, int* __restrict__ best /* 2 */,
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
                const double onep = __dadd_rn(1.0, fwd_factor[ba]);
                if (bc < L - 1 && ba * 15 < 180 && ba > 0) prev[(size_t)(bc + 1) * S] = __dmul_rn(prev[(size_t)(bc + 1) * S], onep);
                else if (bc > 0 && ba * 15 > 180 && ba * 15 < 360) prev[(size_t)(bc - 1) * S] = __dmul_rn(prev[(size_t)(bc - 1) * S], onep);
            }
            __syncthreads();
        }
    }
    // ---- prevWeight then probUpdate
    const double cw1 = 0.7, pw1 = 1 - 0.7;
    const double cw2 = (blur > 200) ? 0.85 : __dmul_rn(__ddiv_rn(blur, 200.0), 0.85);
    const double pw2 = __dsub_rn(1.0, cw2);
    if (stages & 6) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const double p = prev[i];
            double x = cur[i];
            if (stages & 2) x = __dadd_rn(__dmul_rn(cw1, x), __dmul_rn(pw1, p));
            if (stages & 4) x = __dadd_rn(__dmul_rn(cw2, x), __dmul_rn(pw2, p));
            out[i] = x;
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
