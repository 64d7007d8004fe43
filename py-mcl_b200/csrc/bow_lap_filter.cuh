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
    unsigned long long* __restrict__ best /* [n_loc][n_q] */) {
    if (only_if && *only_if == 0) return;
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
// The query vector has at most (rows of the frame) non-zeros out of W = 10000 words, so after normalisation it is
// compacted, in word order (deterministic summation order), into a (word, value) list and every image row is only
// gathered at those words: ~300 x 100 gathers per location instead of a dense 4 MB read.
// grid = (n_locations, n_frames), block = 256; dynamic smem = W floats + list_cap x (int + float).
__global__ void __launch_bounds__(256) bow_score_kernel(
    const unsigned long long* __restrict__ best /* [n_loc][n_q] */, int n_q, const int64_t* __restrict__ q_off /* frames */,
    const float* __restrict__ idf /* [n_loc][W] */, const float* __restrict__ im_features /* rows of W */,
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
    const int64_t i0 = img_off[loc], i1 = img_off[loc + 1];
    for (int64_t im = i0 + warp; im < i1; im += 8) {
        const float* row = im_features + im * W;
        float acc = 0.f;
        for (int k = lane; k < n_nz; k += 32) acc = fmaf(nz_val[k], row[nz_word[k]], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) out_scores[(int64_t)frame * total_images + im] = acc;
    }
}

// ------------------------------------------------------------------------------------------
// variance of the Laplacian (analyze.py:441-445): 4-neighbour kernel, BORDER_REFLECT_101, population var.
// HBM-bound by nature (one byte per pixel).  Each thread owns a 16-pixel-wide column block over LAP_RY rows and slides a
// three-row window down it: one aligned 128-bit load per row (rows are stored with a 16-byte pitch), each row unpacked
// to ints once and used three times from registers; ~6 integer instructions per pixel.  Exact integer sums
// (sum L in s64, sum L^2 in u64; per-thread partials fit 32 bits).
// grid = (ceil(W/1024), ceil(H/(4*LAP_RY)), n_images), block = (64, 4).
// ------------------------------------------------------------------------------------------
constexpr int LAP_RY = 16;

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

struct LapRow {
    int px[16];
    int lh, rh;  // the pixels left of px[0] and right of px[15] (already border-reflected)
};

__device__ __forceinline__ void lap_load_row(const uint8_t* __restrict__ img, int pitch, int W, int y, int x0, LapRow& r) {
    const uint8_t* row = img + (int64_t)y * pitch;
    const uint4 v = *reinterpret_cast<const uint4*>(row + x0);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 16; ++i) r.px[i] = (int)((w[i >> 2] >> (8 * (i & 3))) & 0xffu);
    r.lh = x0 > 0 ? (int)row[x0 - 1] : (W > 1 ? r.px[1] : r.px[0]);
    r.rh = x0 + 16 < W ? (int)row[x0 + 16] : 0;  // only read when the block's last pixel has an in-image right neighbour
}

__global__ void __launch_bounds__(256) laplacian_sums_kernel(const uint8_t* __restrict__ pixels,
                                                             const int64_t* __restrict__ img_off,
                                                             const int* __restrict__ pitches,
                                                             const int* __restrict__ hs, const int* __restrict__ ws,
                                                             long long* __restrict__ sums /* [n][2] */) {
    const int im = blockIdx.z;
    const int H = hs[im], W = ws[im], pitch = pitches[im];
    const int x0 = (blockIdx.x * 64 + threadIdx.x) * 16;
    const int ys = (blockIdx.y * 4 + threadIdx.y) * LAP_RY;
    int s1 = 0;
    unsigned s2 = 0;
    if (x0 < W && ys < H) {
        const uint8_t* img = pixels + img_off[im];
        const int ye = min(H, ys + LAP_RY);
        LapRow up, mid, dn;
        lap_load_row(img, pitch, W, reflect101(ys - 1, H), x0, up);
        lap_load_row(img, pitch, W, ys, x0, mid);
        for (int y = ys; y < ye; ++y) {
            lap_load_row(img, pitch, W, reflect101(y + 1, H), x0, dn);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int x = x0 + i;
                const int left = i == 0 ? mid.lh : mid.px[i - 1];
                int right = i == 15 ? mid.rh : mid.px[i + 1];
                if (x == W - 1) right = W > 1 ? left : mid.px[i];  // REFLECT_101: the neighbour beyond the edge mirrors x-1
                const int l = up.px[i] + dn.px[i] + left + right - 4 * mid.px[i];
                if (x < W) {
                    s1 += l;
                    s2 += (unsigned)(l * l);
                }
            }
            up = mid;
            mid = dn;
        }
    }
    long long t1 = s1;
    unsigned long long t2 = s2;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        t1 += __shfl_xor_sync(0xffffffffu, t1, o);
        t2 += __shfl_xor_sync(0xffffffffu, t2, o);
    }
    if (threadIdx.x % 32 == 0 && (t1 != 0 || t2 != 0)) {
        atomicAdd(reinterpret_cast<unsigned long long*>(&sums[2 * im]), (unsigned long long)t1);
        atomicAdd(reinterpret_cast<unsigned long long*>(&sums[2 * im + 1]), t2);
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

__global__ void filter_step_kernel(int L, int I, int stages, int cmd, double blur, int blur_t, double* __restrict__ prev,
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
                if (bc < L - 1 && ba * 15 < 180 && ba > 0) row = bc + 1;
                else if (bc > 0 && ba * 15 > 180 && ba * 15 < 360) row = bc - 1;
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
