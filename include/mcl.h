/* libmcl — C-ABI of the B200-native descriptor-matching hot path of zhangxingshuo/py-mcl.
 *
 * The reference exposes no FFI: its seam is the Python object API (Matcher.py:35-388, analyze.py:32-445)
 * on top of the implicit cv2 / scipy matcher contracts.  Each entry point below names the reference
 * interface it replaces (file:line relative to the reference checkout).  The only caller is the ctypes
 * shim in py-mcl_b200/_lib.py; INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success and a negative MCL_E* code on error; mcl_last_error() returns a
 *     thread-local message for the last failure on the calling thread.
 *   - *_device functions use the context's scratch buffers (candidate lists, staging): calls on one context must be
 *     issued from one host thread and are ordered like their streams; do not run two of them concurrently on
 *     different streams of the same context.
 *   - host buffers are caller-owned and only read/written during the call; device memory is owned by handles.
 *   - one host thread per handle.  Functions without a `stream` argument are synchronous on return;
 *     *_device functions only enqueue work on `stream` (a cudaStream_t, NULL = the context's own stream).
 *   - there is NO CPU fallback: mcl_init fails unless the device is compute capability 10.x (sm_100a code).
 */
#ifndef MCL_H_
#define MCL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mcl_ctx mcl_ctx;
typedef struct mcl_map mcl_map;
typedef struct mcl_queries mcl_queries;
typedef struct mcl_bow mcl_bow;

enum { MCL_OK = 0, MCL_EINVAL = -1, MCL_ECUDA = -2, MCL_ENODEV = -3, MCL_EDATA = -4, MCL_EINTERNAL = -5 };
/* descriptor kinds: extractor output of Matcher.createFeatureIndex (Matcher.py:147-163) */
enum { MCL_SIFT = 0 /* 128 x u8 (integer-valued f32 accepted) */, MCL_ORB = 1 /* 32 x u8 */,
       MCL_SURF = 2 /* 64 x f32 */, MCL_BOWQ = 3 /* 128 x f32 query rows for mcl_bow_* */ };
enum { MCL_U8 = 0, MCL_F32 = 1 };
/* match modes */
enum { MCL_KNN2_RATIO = 0 /* knnMatch(k=2) + ratio lambda, Matcher.py:218-219,279-280 */,
       MCL_CROSSCHECK = 1 /* ORB only: BFMatcher(NORM_HAMMING, crossCheck=True).match, Matcher.py:180-182 */ };
/* algorithm selector for SIFT (testing / profiling) */
enum { MCL_ALGO_AUTO = 0, MCL_ALGO_TENSOR = 1 /* tcgen05: SIFT sift_tc4 (A in tensor memory, norm-sorted groups, N = 128); ORB knn on
                                                 the tensor cores (tc9, +-1 rows) */,
       MCL_ALGO_SIMT = 2 /* SIFT dp4a / ORB xor+popc, exact top-2 per thread */ };

/* ---- context ---------------------------------------------------------------------------------------- */
int mcl_init(int device_id, mcl_ctx** out);
void mcl_shutdown(mcl_ctx* ctx);
const char* mcl_last_error(void);
/* info[0]=SM count, [1]=cc major, [2]=cc minor, [3]=HBM MiB, [4]=max SM clock kHz, [5]=L2 bytes, [6]=opt-in smem bytes,
 * [7]=whole-image exact rescans done by the SIFT fix-up kernels so far (statistics) */
int mcl_device_info(mcl_ctx* ctx, int64_t info[8]);
int mcl_sync(mcl_ctx* ctx);
/* statistics since mcl_init (synchronises): out[0] = SIFT candidates resolved by the exact fix-up, [1] = SIFT whole-image exact
 * rescans, [2] = SIFT/ORB (frame, image) pairs recomputed by the SIMT kernel because a candidate list was full, [3] = BOW
 * rows rescanned over a whole code book, [4] = ORB candidates resolved exactly, [6] = entries of the BOW resolve list in the
 * last call, [5], [7] reserved.
 * No reference counterpart (bench.py reports them per input distribution). */
int mcl_stats(mcl_ctx* ctx, int64_t out[8]);
/* pinned host memory for the end-to-end path (cudaHostAlloc) */
int mcl_host_alloc(size_t bytes, void** out);
void mcl_host_free(void* p);
/* tuning / diagnostic switches (development aid, no reference counterpart).  option 0: SIFT tensor-core kernel
 * epilogue variant, 0 = product, other values = profiling variants with meaningless counts; option 1 = 1: tile records from
 * global memory; option 2 = 1: BOW word assignment on the float32 SIMT kernel only; option 3 = n > 0: cap the SIFT / ORB
 * candidate list at n entries (tests: forces the dirty-pair -> exact SIMT repair path); option 5: thread-block clusters of 2
 * sharing the map tiles by multicast, 0 = automatic, 1 = never, 2 = always; option 6 = n > 0: degrees between neighbouring map
 * images for mcl_filter_step's forward-command rule (the reference's 15, analyze.py:331-335).
 * The environment variable MCL_OPT="option=value,..." presets these at mcl_init (experiments). */
int mcl_set_option(mcl_ctx* ctx, int option, int value);
/* number of kernels this library launched since mcl_init (bench.py's gpu_launches) */
int64_t mcl_launch_count(mcl_ctx* ctx);
/* per-kernel device timing with CUDA events on the launching stream.  which: 0 = SIFT tensor-core kernel,
 * 1 = SIFT fix-up, 2 = ORB, 3 = SURF, 4 = BOW vq, 5 = BOW score, 6 = Laplacian, 7 = ingest (prep),
 * 8 = SIFT SIMT (dp4a) kernel, 9 = BOW vq tensor-core kernel, 10 = BOW float32 resolve kernels, 11 = chi-squared kernel,
 * 12 = ORB exact fix-up / repair kernels (slot 2 = the ORB tensor-core or SIMT matching kernel), 13 = BOW whole-code-book
 * rescans of ambiguous rows.
 * mcl_profile_read synchronises, returns accumulated milliseconds and launch count, and resets them. */
int mcl_profile_enable(mcl_ctx* ctx, int on);
int mcl_profile_read(mcl_ctx* ctx, int which, double* ms, int64_t* launches);

/* ---- map index ---------------------------------------------------------------------------------------
 * replaces the per-location dict built by Matcher.createFeatureIndex (Matcher.py:147-163) and consumed through
 * self.index[imagePath] (Matcher.py:178,210,271): all images of this rank's shard, concatenated.
 * row_offsets has n_images+1 entries; desc is host memory, row-major (rows, 128|32|64) of desc_dtype.
 * The descriptors stay resident in HBM (SIFT: u8 in the tcgen05 SWIZZLE_128B image + int32 norms). */
int mcl_map_create(mcl_ctx* ctx, int kind, int n_images, const int64_t* row_offsets, const void* desc,
                   int desc_dtype, mcl_map** out);
void mcl_map_destroy(mcl_map* map);
/* info[0]=n_images, [1]=rows, [2]=padded rows, [3]=resident bytes */
int mcl_map_info(mcl_map* map, int64_t info[4]);

/* ---- query batch ---------------------------------------------------------------------------------------
 * replaces des1 of *Match (Matcher.py:177,205,242,270), extracted ONCE per frame for a batch of
 * n_queries frames; q_row_offsets has n_queries+1 entries. */
int mcl_queries_create(mcl_ctx* ctx, int kind, int n_queries, const int64_t* q_row_offsets, const void* q_desc,
                       int q_dtype, mcl_queries** out);
/* The same batch built from rows that are ALREADY in device memory (d_desc: rows x 128|32|64 of q_dtype, row-major;
 * q_row_offsets stays a host array).  Only enqueues work on `stream` (NULL = the context's stream); the multi-GPU driver
 * uploads each frame on ONE rank and all-gathers the rows over NVLink, so every rank builds its batch from device memory
 * (replaces the same des1 as above).  Integer-valuedness of float32 SIFT rows is recorded on the device:
 * mcl_queries_validate synchronises and returns MCL_EDATA if the ingest saw a non-integer value. */
int mcl_queries_create_device(mcl_ctx* ctx, int kind, int n_queries, const int64_t* q_row_offsets, const void* d_desc,
                              int q_dtype, void* stream, mcl_queries** out);
int mcl_queries_validate(mcl_queries* q);
void mcl_queries_destroy(mcl_queries* q);

/* ---- matching ------------------------------------------------------------------------------------------
 * out_counts[q*n_images + i] = len(good) of SIFTMatch/SURFMatch/ORBMatch(query q, image i)
 * (Matcher.py:169-194, Matcher.py:196-230, Matcher.py:262-291), i.e. what Matcher.run's loop (Matcher.py:309-318) collects.
 * mask (optional, n_queries*n_images bytes): 0 = image not matched for that query -> fill_value
 * (optRun's "numMatches = 1", Matcher.py:359,369).  Images with fewer than 2 descriptors count 0.
 * mcl_match_counts: host in / host out, includes H2D, ingest, kernels, D2H.
 * mcl_match_counts_device: resident queries, device mask / counts pointers, asynchronous on `stream`. */
int mcl_match_counts(mcl_map* map, int n_queries, const int64_t* q_row_offsets, const void* q_desc, int q_dtype,
                     int mode, double ratio, const uint8_t* mask, int32_t fill_value, int algo,
                     int32_t* out_counts);
int mcl_match_counts_device(mcl_map* map, mcl_queries* q, int mode, double ratio, const uint8_t* d_mask,
                            int32_t fill_value, int algo, int32_t* d_counts, void* stream);

/* ---- bag of words --------------------------------------------------------------------------------------
 * replaces Matcher.BOWMatch (Matcher.py:232-259) for n_locations indices at once: per location the tuple
 * (im_features, idf, voc) of joblib.load(indexPath) (Matcher.py:236), kept resident instead of re-loaded.
 * voc_row_offsets / img_offsets have n_locations+1 entries; idf is (n_locations, W); im_features is
 * (total_images, W).  out_words (optional) is (n_locations, total query rows) int32 = vq()'s `words`
 * (Matcher.py:250); out_scores is (n_queries, total_images) float32 = `score` (Matcher.py:258). */
int mcl_bow_create(mcl_ctx* ctx, int n_locations, const int64_t* voc_row_offsets, const float* voc,
                   const float* idf, const int64_t* img_offsets, const float* im_features, int W, mcl_bow** out);
void mcl_bow_destroy(mcl_bow* bow);
int mcl_bow_score(mcl_bow* bow, int n_queries, const int64_t* q_row_offsets, const void* q_desc, int q_dtype,
                  int32_t* out_words, float* out_scores);
int mcl_bow_score_device(mcl_bow* bow, mcl_queries* q, int32_t* d_words, float* d_scores, void* stream);

/* mcl_kmeans: the code book training of Matcher.createIndex (Matcher.py:124, `kmeans(descriptors, numWords, 1)`), i.e.
 * scipy.cluster.vq._kmeans(obs, guess, thresh): Lloyd iterations from the initial code book `guess` (k x 128 float32) over
 * n observations (n x 128, uint8 or integer-valued float32) until the average distortion changes by <= thresh; empty
 * clusters are dropped.  out_voc must hold k x 128 floats; *out_k = rows kept.  The assignment step is mcl_bow_score's. */
int mcl_kmeans(mcl_ctx* ctx, int64_t n, const void* obs, int obs_dtype, int k, const float* guess, double thresh,
               int max_iter, float* out_voc, int* out_k, double* out_distortion, int* out_iters);

/* ---- small stages --------------------------------------------------------------------------------------
 * mcl_laplacian_var: analyzer.Laplacian (analyze.py:441-445) for n grayscale u8 images (host pointers,
 * row stride in bytes): out[i] = cv2.Laplacian(img, CV_64F).var(), bit for bit: numpy's two-pass variance with its pairwise
 * summation order (csrc/lap_var.cuh).
 * mcl_laplacian_sums: the exact integer sums behind it, sums[2i] = sum L, sums[2i+1] = sum L^2. */
int mcl_laplacian_var(mcl_ctx* ctx, int n, const uint8_t* const* imgs, const int* h, const int* w,
                      const int* stride, double* out);
int mcl_laplacian_sums(mcl_ctx* ctx, int n, const uint8_t* const* imgs, const int* h, const int* w,
                       const int* stride, int64_t* sums);
/* mcl_laplacian_var for images the caller has already packed the way the device keeps them -- every row padded to a
 * multiple of 16 bytes, image i at byte offset sum_{j<i} h[j] * ((w[j] + 15) & ~15) -- in ONE (ideally pinned, mcl_host_alloc)
 * host buffer: one host-to-device copy instead of one strided copy per image (analyze.py:441-445 for a whole sequence). */
int mcl_laplacian_var_packed(mcl_ctx* ctx, int n, const uint8_t* packed, const int* h, const int* w, double* out);
/* mcl_filter_step: one frame of analyzer.processRaw (analyze.py:119-134).  `stages` selects which of the
 * reference's steps run, in this order:
 *   MCL_F_COMMAND  accountCommand (analyze.py:312-336) on `prev`, in place like the reference's shallow copy
 *   MCL_F_PREV     prevWeight (analyze.py:278-310):  out = 0.7*cur + (1-0.7)*prev
 *   MCL_F_BLUR     probUpdate (analyze.py:239-276):  out = cw*X + (1-cw)*prev, X = the prevWeight result when
 *                  MCL_F_PREV is also set, else cur;  cw = 0.85 if blur > 200 else blur/200*0.85
 *   MCL_F_BEST     best[0] = index of the lexicographic max row of out, best[1] = first argmax of its
 *                  probabilities (analyze.py:133-134)
 * prev, cur, out: L rows of [total, p_0..p_{I-1}] float64, arithmetic in float64 without FMA contraction. */
enum { MCL_F_COMMAND = 1, MCL_F_PREV = 2, MCL_F_BLUR = 4, MCL_F_BEST = 8, MCL_F_ALL = 15 };
int mcl_filter_step(mcl_ctx* ctx, int L, int I, int stages, char cmd, double blur, double* prev, const double* cur,
                    double* out, int* best);

/* mcl_filter_step_typed: the same step when the lists hold numpy scalars as well as Python numbers, which is what the
 * colour method produces (np.float32 from the float32 histograms; np.float64 from ndarray.var() in the blur weight,
 * analyze.py:245-248 + :441-445).  Python's operators then follow numpy's scalar promotion -- float32 with a Python
 * float stays float32 (the Python value is rounded to float32 first), anything with np.float64 is float64 -- and the
 * kernel applies that rule to every multiply and add.  Values travel as float64 (float32 values are exact in it);
 * prev_type / cur_type / out_type hold two tags per location row, [total, probabilities]; blur_type is the tag of `blur` (MCL_T_F64 for the
 * reference's Laplacian); NULL type arrays mean MCL_T_PY, with which the result equals mcl_filter_step's. */
enum { MCL_T_PY = 0, MCL_T_F64 = 1, MCL_T_F32 = 2 };
int mcl_filter_step_typed(mcl_ctx* ctx, int L, int I, int stages, char cmd, double blur, int blur_type, double* prev,
                          const int8_t* prev_type, const double* cur, const int8_t* cur_type, double* out,
                          int8_t* out_type, int* best);

/* mcl_chi2: the colour-histogram search (search.py:7-26 Searcher.search / chisquared, called from Matcher.colorSearch,
 * Matcher.py:89-99): out[q * n_map + m] = 0.5 * sum_b (a_b - b_b)^2 / (a_b + b_b + eps) for every (query histogram,
 * map histogram) pair, host pointers, row-major (n, bins) float32.  Arithmetic is float32 in numpy's pairwise-summation
 * order, i.e. bit-identical to the reference's np.float32 results on float32 histograms (cv2.calcHist + normalize);
 * the reference's eps is 1e-10.  bins = 512 (the reference's [8, 8, 8]) takes the warp-per-pair kernel. */
int mcl_chi2(mcl_ctx* ctx, int n_q, int n_map, int bins, const float* q_hist, const float* map_hist, float eps,
             float* out);

/* ---- roofline denominators ------------------------------------------------------------------------------
 * MEASURED_PEAKS.json carries HBM GB/s and bf16 TFLOP/s only; the hot path is bounded by the int8 tensor pipe
 * (SIFT) and the popc pipe (ORB).  which: 0 = tcgen05 kind::i8 128x128x32 MMAs issued back to back from shared
 * memory on every SM (out[0] = int8 ops/s, 2 per MAC), 1 = XOR+popc.b32 issue rate (out[0] = popc32 results/s), 2 = fp32
 * (a - b)^2 accumulation, one FADD + one FFMA per element as in SURF's direct-difference distance (out[0] = elements/s).
 * out[1] = elapsed milliseconds.  No reference counterpart (SURVEY.md section 8d asks for these). */
int mcl_microbench(mcl_ctx* ctx, int which, int iters, double* out);

#ifdef __cplusplus
}
#endif
#endif /* MCL_H_ */
