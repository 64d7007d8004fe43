// libmcl.so — host side of the C-ABI declared in include/mcl.h (sm_100a only, no CPU fallback).
#include "../../include/mcl.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "prep.cuh"
#include "sift_tc3.cuh"
#include "sift_tc4.cuh"
#include "simt_match.cuh"
#include "bow_lap_filter.cuh"
#include "lap_var.cuh"
#include "tc9.cuh"
#include "bow_tc.cuh"
#include "kmeans.cuh"
#include "color.cuh"
#include "microbench.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(expr)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (expr);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(MCL_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

enum { PROF_SIFT_TC = 0, PROF_SIFT_FIX, PROF_ORB, PROF_SURF, PROF_VQ, PROF_SCORE, PROF_LAP, PROF_PREP, PROF_SIFT_SIMT, PROF_VQ_TC, PROF_VQ_RESOLVE, PROF_CHI2, PROF_ORB_FIX, PROF_VQ_RESCAN, PROF_N };
constexpr int QROW_PAD = 384;  // sift_tc4's query block

// growable device scratch buffer (never shrinks; avoids cudaMalloc on the per-frame path)
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = std::max(bytes, (size_t)4096);
        want = (want + 255) & ~(size_t)255;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace

// numpy's pairwise summation of N elements as tables (lap_var.cuh): the leaves and, level by level, the internal nodes
struct LapPlan {
    int n_leaves = 0, n_nodes = 0, n_levels = 0;
    int* d_tab = nullptr;   // leaf_lo[n_leaves] | leaf_n[n_leaves] | node_l[n_nodes] | node_r[n_nodes] | level_off[n_levels + 1]
};

struct mcl_ctx {
    int device = 0;
    int sm_count = 0;
    cudaDeviceProp prop{};
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // H2D + ingest of the next part of a host batch while the current part is matched
    std::vector<cudaEvent_t> part_events;
    int64_t launches = 0;
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof[PROF_N];
    // scratch
    DevBuf diag9;        // MCL_TC9_DIAG
    DevBuf dirty_list;   // (frame * n_images + image) of the dirty pairs, compacted on the device
    DevBuf stage, cand, dirty, flags /* ints: [0]=cand_count [1]=overflow [2]=bad [3]=rescans [4]=dirty pairs listed [6]=bow bad [8..10]=bow state; bytes 64..: statistics */, row_arg, bow_best, lap_pix, lap_meta, filt, tmp_counts,
        tmp_mask, tmp_words, tmp_scores, chi2, tq_off, tq_desc, tq_norm, tq_frame, bow_q8, bow_qaux, bow_tc_out, bow_sort, orb_q;
    int fixup_blocks = 0;
    bool tc9_attr = false;
    int tc9_clusters[2] = {0, 0};   // co-resident tc9 CTA pairs (ORB, BOW)
    size_t cand_cap_entries = 0;
    bool tc4_attr = false;
    int opt[8] = {0};  // mcl_set_option: [0] = sift_tc4 diagnostic mode
    int max_clusters2 = 0;          // co-resident sift_tc4 clusters of 2 CTAs
    int bow_score_smem = 0;         // dynamic shared memory opt-in of bow_score_kernel so far
    int64_t active_pairs_hint = -1; // (frame, image) pairs a host DOR mask leaves active, for the launch shape of small problems
    std::map<int64_t, LapPlan> lap_plans;   // by pixel count
    int lap_smem_optin = 48 * 1024;
    // free list of device blocks for the query batches of mcl_queries_create_device: the per-step batches of the multi-GPU
    // driver come and go with the same few sizes, and cudaMalloc / cudaFree (which synchronises the device) cost more than
    // the ingest itself
    std::vector<std::pair<void*, size_t>> pool;
    size_t pool_bytes = 0;
    DevBuf lap_vals, lap_list;
};

struct mcl_queries {
    mcl_ctx* ctx = nullptr;
    int kind = 0;
    int n_queries = 0;
    int64_t rows = 0, padded_rows = 0;
    std::vector<int64_t> h_off;
    int64_t* d_off = nullptr;      // n_queries+1 (source rows == resident rows for queries)
    void* d_desc = nullptr;        // SIFT: SW128 atoms; ORB: (rows,32) u8; SURF: (rows,64) f32; BOWQ: (rows,128) f32
    int32_t* d_norm = nullptr;     // SIFT
    int32_t* d_frame = nullptr;    // SIFT: frame of each resident row
    int* d_bad = nullptr;          // device-built batches: set by the ingest kernel when a float row is not integer valued (mcl_queries_validate)
    cudaStream_t built_on = nullptr;
    bool pooled = false;           // device blocks come from / go back to the context's free list
    size_t sz_off = 0, sz_desc = 0, sz_norm = 0, sz_frame = 0;
    bool borrowed = false;         // buffers belong to ctx scratch (temporary batch of mcl_match_counts)
};

struct mcl_map {
    mcl_ctx* ctx = nullptr;
    int kind = 0;
    int n_images = 0;
    int64_t rows = 0, padded_rows = 0;
    size_t resident_bytes = 0;
    std::vector<int64_t> h_src_off, h_dst_off;
    int64_t* d_src_off = nullptr;
    int64_t* d_dst_off = nullptr;
    void* d_desc = nullptr;
    int32_t* d_norm = nullptr;
    int32_t* d_img_of = nullptr;
    int n_tiles = 0;                     // 128-row tiles
    // chunk tables for a few granularities (128-row tiles per chunk), picked per call from the query batch size
    struct ChunkSet { int target = 0; int n = 0; mcl::sifttc3::Chunk* d = nullptr; };
    std::vector<ChunkSet> chunk_sets;
    mcl::sifttc3::TileMeta* d_meta3 = nullptr;   // per 64-row unit: image ids, descriptor counts, norm brackets
    int n_units = 0;                             // 64-row units
    // ORB on the tensor cores (tc9.cuh): +-1 byte tiles, the packed rows in padded order, tile records, rows per image
    uint8_t* d_tiles9 = nullptr;
    uint8_t* d_packed9 = nullptr;
    mcl::tc9::TileRec* d_rec9 = nullptr;
    int32_t* d_seg_rows = nullptr;
    int max_rows_per_image = 0;
};

struct mcl_bow {
    mcl_ctx* ctx = nullptr;
    int n_loc = 0, W = 0;
    int64_t total_images = 0, voc_rows = 0;
    int max_words = 0;
    std::vector<int64_t> h_voc_off, h_img_off;
    // tensor-core word assignment (tc9.cuh, KIND_BOW): fp16 code-book tiles with the norm extension, per 128 words
    std::vector<int32_t> h_tile_off;
    int32_t* d_tile_off = nullptr;
    uint8_t* d_tiles = nullptr;
    mcl::tc9::TileRec* d_rec = nullptr;
    float* d_key_c = nullptr;
    struct Split { int n_splits = 0, n_chunks = 0; mcl::sifttc3::Chunk* d = nullptr; };
    std::vector<Split> splits;   // chunk tables: every location's tiles cut into 1, 2, 4, 8 ranges
    int cap_tiles = 0, cap_loc = 0;   // what the tensor-core buffers were allocated for (k-means rebuilds them every iteration)
    bool tc_ok = false;
    float* d_voc = nullptr;
    float* d_idf = nullptr;
    float* d_imf = nullptr;
    int64_t* d_voc_off = nullptr;
    int64_t* d_img_off = nullptr;
};

namespace {

cudaError_t pool_alloc(mcl_ctx* c, void** out, size_t bytes) {
    bytes = (std::max<size_t>(bytes, 256) + 255) & ~(size_t)255;
    int best = -1;
    for (int i = 0; i < (int)c->pool.size(); ++i)
        if (c->pool[i].second >= bytes && c->pool[i].second <= 2 * bytes + (1 << 20) && (best < 0 || c->pool[i].second < c->pool[best].second)) best = i;
    if (best >= 0) {
        *out = c->pool[best].first;
        c->pool_bytes -= c->pool[best].second;
        c->pool.erase(c->pool.begin() + best);
        return cudaSuccess;
    }
    return cudaMalloc(out, bytes);
}
void pool_free(mcl_ctx* c, void* p, size_t bytes) {
    if (!p) return;
    bytes = (std::max<size_t>(bytes, 256) + 255) & ~(size_t)255;
    if (c->pool.size() >= 64 || c->pool_bytes + bytes > ((size_t)4 << 30)) { cudaFree(p); return; }
    c->pool.push_back({p, bytes});
    c->pool_bytes += bytes;
}

struct ProfScope {
    mcl_ctx* c;
    int which;
    cudaStream_t st;
    cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(mcl_ctx* c_, int w, cudaStream_t s) : c(c_), which(w), st(s) {
        if (c->profiling) {
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            cudaEventRecord(a, st);
        }
        c->launches++;
    }
    ~ProfScope() {
        if (a) {
            cudaEventRecord(b, st);
            c->prof[which].push_back({a, b});
        }
    }
};

inline int grid_for(int64_t n, int block, int cap) {
    int64_t g = (n + block - 1) / block;
    return (int)std::max<int64_t>(1, std::min<int64_t>(g, cap));
}

int check_offsets(const int64_t* off, int n, const char* what) {
    if (!off) return fail(MCL_EINVAL, "%s: offsets pointer is NULL", what);
    if (off[0] != 0) return fail(MCL_EINVAL, "%s: offsets[0] must be 0", what);
    for (int i = 0; i < n; ++i)
        if (off[i + 1] < off[i]) return fail(MCL_EINVAL, "%s: offsets must be non-decreasing", what);
    return MCL_OK;
}

size_t row_bytes(int kind, int dtype) {
    switch (kind) {
        case MCL_SIFT: return dtype == MCL_F32 ? 512 : 128;
        case MCL_ORB: return 32;
        case MCL_SURF: return 256;
        case MCL_BOWQ: return dtype == MCL_F32 ? 512 : 128;
    }
    return 0;
}

// ingest (n_rows,128) rows from the host into the SW128-atom layout + norms + segment ids
int ingest_rows128(mcl_ctx* c, cudaStream_t st, const void* h_src, int dtype, int64_t n_rows, int n_seg,
                   const int64_t* d_src_off, const int64_t* d_dst_off, int64_t padded_rows, uint8_t* d_desc,
                   int32_t* d_norm, int32_t* d_seg, uint8_t* d_stage = nullptr /* null: the context's staging buffer */,
                   bool check_now = true, const int32_t* d_perm = nullptr, cudaStream_t copy_st = nullptr,
                   cudaEvent_t copied = nullptr, bool src_on_device = false, int* d_bad = nullptr) {
    CU(cudaMemsetAsync(d_desc, 0, (size_t)padded_rows * 128, st));
    CU(cudaMemsetAsync(d_norm, 0xFF, (size_t)padded_rows * 4, st));
    CU(cudaMemsetAsync(d_seg, 0xFF, (size_t)padded_rows * 4, st));
    if (n_rows == 0) return MCL_OK;
    const size_t bytes = (size_t)n_rows * (dtype == MCL_F32 ? 512 : 128);
    if (src_on_device) {
        d_stage = (uint8_t*)const_cast<void*>(h_src);   // rows are device memory already: ingest straight from them
    } else if (!d_stage) {
        CU(c->stage.ensure(bytes));
        d_stage = c->stage.as<uint8_t>();
    }
    if (src_on_device) {
    } else if (copy_st) {
        // the H2D copy goes on a stream that carries nothing but copies: a memset or conversion kernel queued in front of
        // it would wait for SMs behind the persistent matching kernel and hold the copy back with it
        CU(cudaMemcpyAsync(d_stage, h_src, bytes, cudaMemcpyHostToDevice, copy_st));
        CU(cudaEventRecord(copied, copy_st));
        CU(cudaStreamWaitEvent(st, copied, 0));
    } else {
        CU(cudaMemcpyAsync(d_stage, h_src, bytes, cudaMemcpyHostToDevice, st));
    }
    int* bad = d_bad ? d_bad : c->flags.as<int>() + 2;
    if (check_now) CU(cudaMemsetAsync(bad, 0, 4, st));
    {
        ProfScope ps(c, PROF_PREP, st);
        const int grid = grid_for(n_rows * 32, 256, c->sm_count * 8);
        if (dtype == MCL_F32)
            mcl::prep_rows128_kernel<true><<<grid, 256, 0, st>>>(d_stage, n_rows, d_src_off, d_dst_off, n_seg, d_desc, d_norm, d_seg, bad, d_perm);
        else
            mcl::prep_rows128_kernel<false><<<grid, 256, 0, st>>>(d_stage, n_rows, d_src_off, d_dst_off, n_seg, d_desc, d_norm, d_seg, bad, d_perm);
    }
    CU(cudaGetLastError());
    if (dtype == MCL_F32 && check_now) {
        int h_bad = 0;
        CU(cudaMemcpyAsync(&h_bad, bad, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (h_bad) return fail(MCL_EDATA, "SIFT descriptors must be integer valued in [0,255]");
    }
    return MCL_OK;
}

void build_chunks(const mcl_map* m, int TILE_N, int target_tiles, std::vector<mcl::sifttc3::Chunk>& out) {
    out.clear();
    int i = 0;
    const int n = m->n_images;
    while (i < n) {
        // skip empty images (they never get a group; their count stays 0)
        if (m->h_dst_off[i + 1] == m->h_dst_off[i]) { ++i; continue; }
        mcl::sifttc3::Chunk ch;
        ch.img_begin = i;
        ch.tile_begin = (int)(m->h_dst_off[i] / TILE_N);
        int j = i;
        int tile_end = ch.tile_begin;
        while (j < n) {
            if (m->h_dst_off[j + 1] > m->h_dst_off[j]) tile_end = (int)((m->h_dst_off[j + 1] + TILE_N - 1) / TILE_N);
            ++j;
            if (tile_end - ch.tile_begin >= target_tiles) break;
        }
        ch.img_end = j;
        ch.tile_end = tile_end;
        out.push_back(ch);
        i = j;
    }
}

const mcl_map::ChunkSet* pick_chunks(const std::vector<mcl_map::ChunkSet>& sets, int n_qblocks, int sm_count, int64_t total_tiles,
                                     double refill = 2.5) {
    // every work item pays a pipeline refill (the query tiles in tensor memory are single buffered: ~2.5 tile times), and
    // the last round of items is only partly filled: take the set with the smallest estimated makespan
    const mcl_map::ChunkSet* best = &sets.front();
    double best_cost = 1e300;
    for (const auto& cs : sets) {
        if (cs.n == 0) continue;
        const int64_t items = (int64_t)cs.n * n_qblocks;
        const int64_t rounds = (items + sm_count - 1) / sm_count;
        const double cost = (double)rounds * ((double)total_tiles / cs.n + refill);
        if (cost < best_cost) { best_cost = cost; best = &cs; }
    }
    return best;
}

}  // namespace

// ================================================================================================
// context
// ================================================================================================
extern "C" const char* mcl_last_error(void) { return g_err.c_str(); }

extern "C" int mcl_init(int device_id, mcl_ctx** out) {
    if (!out) return fail(MCL_EINVAL, "mcl_init: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(MCL_ENODEV, "no CUDA device (%s); libmcl has no CPU fallback", e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
    if (device_id < 0 || device_id >= n) return fail(MCL_EINVAL, "device %d out of range (have %d)", device_id, n);
    CU(cudaSetDevice(device_id));
    mcl_ctx* c = new (std::nothrow) mcl_ctx();
    if (!c) return fail(MCL_EINTERNAL, "out of host memory");
    c->device = device_id;
    e = cudaGetDeviceProperties(&c->prop, device_id);
    if (e != cudaSuccess) { delete c; return fail(MCL_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); }
    if (c->prop.major != 10) {
        const int mj = c->prop.major, mn = c->prop.minor;
        delete c;
        return fail(MCL_ENODEV, "device %d is sm_%d%d; libmcl is built for sm_100a only and has no fallback", device_id, mj, mn);
    }
    c->sm_count = c->prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return fail(MCL_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    e = c->flags.ensure(256);
    if (e != cudaSuccess) { delete c; return fail(MCL_ECUDA, "cudaMalloc: %s", cudaGetErrorString(e)); }
    cudaMemset(c->flags.p, 0, 256);
    if (const char* e = getenv("MCL_OPT")) {   // experiments: "option=value,option=value" presets mcl_set_option
        int o = 0, v = 0, n = 0;
        while (sscanf(e, "%d=%d%n", &o, &v, &n) == 2) {
            if (o >= 0 && o < 8) c->opt[o] = v;
            e += n;
            if (*e == ',') ++e;
        }
    }
    *out = c;
    return MCL_OK;
}

extern "C" void mcl_shutdown(mcl_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto& v : c->prof)
        for (auto& pr : v) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    DevBuf* bufs[] = {&c->stage, &c->cand, &c->dirty, &c->flags, &c->row_arg, &c->bow_best, &c->lap_pix, &c->lap_meta, &c->filt,
                      &c->tmp_counts, &c->tmp_mask, &c->tmp_words, &c->tmp_scores, &c->tq_off, &c->tq_desc, &c->tq_norm,
                      &c->tq_frame, &c->bow_q8, &c->bow_qaux, &c->bow_tc_out, &c->bow_sort, &c->orb_q};
    for (DevBuf* b : bufs) b->release();
    for (auto& kv : c->lap_plans) cudaFree(kv.second.d_tab);
    for (auto& b : c->pool) cudaFree(b.first);
    c->pool.clear();
    c->dirty_list.release();
    c->diag9.release();
    c->lap_vals.release();
    c->lap_list.release();
    for (cudaEvent_t e : c->part_events) cudaEventDestroy(e);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int mcl_device_info(mcl_ctx* c, int64_t info[8]) {
    if (!c || !info) return fail(MCL_EINVAL, "mcl_device_info: NULL argument");
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, c->device);
    info[0] = c->sm_count;
    info[1] = c->prop.major;
    info[2] = c->prop.minor;
    info[3] = (int64_t)(c->prop.totalGlobalMem >> 20);
    info[4] = clk;
    info[5] = c->prop.l2CacheSize;
    info[6] = (int64_t)c->prop.sharedMemPerBlockOptin;
    {   // statistics: exact whole-image rescans done by the SIFT fix-up kernels since mcl_init
        unsigned int r = 0;
        cudaStreamSynchronize(c->stream);
        cudaMemcpy(&r, c->flags.as<unsigned int>() + 3, 4, cudaMemcpyDeviceToHost);
        info[7] = r;
    }
    return MCL_OK;
}

extern "C" int mcl_stats(mcl_ctx* c, int64_t out[8]) {
    if (!c || !out) return fail(MCL_EINVAL, "mcl_stats: NULL argument");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    unsigned int f[16];
    unsigned long long st[4];
    CU(cudaMemcpy(f, c->flags.p, sizeof f, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(st, c->flags.as<unsigned long long>() + 8, sizeof st, cudaMemcpyDeviceToHost));
    out[0] = (int64_t)st[0];   // SIFT candidates resolved by the exact fix-up
    out[1] = f[3];             // SIFT whole-image exact rescans
    out[2] = (int64_t)st[1];   // SIFT (frame, image) pairs recomputed by the SIMT kernel after a full candidate list
    out[3] = f[10];            // BOW rows rescanned over the whole code book
    out[6] = f[11];            // BOW entries of the sorted resolve list in the last call (rows with two candidate groups count twice)
    out[4] = (int64_t)st[2];   // ORB candidates resolved exactly
    out[5] = (int64_t)st[3];
    out[7] = 0;
    return MCL_OK;
}

extern "C" int mcl_sync(mcl_ctx* c) {
    if (!c) return fail(MCL_EINVAL, "mcl_sync: NULL context");
    CU(cudaStreamSynchronize(c->stream));
    return MCL_OK;
}

extern "C" int mcl_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(MCL_EINVAL, "mcl_host_alloc: out is NULL");
    CU(cudaHostAlloc(out, std::max<size_t>(bytes, 1), cudaHostAllocDefault));
    return MCL_OK;
}
extern "C" void mcl_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

extern "C" int mcl_set_option(mcl_ctx* c, int option, int value) {
    if (!c || option < 0 || option >= 8) return fail(MCL_EINVAL, "mcl_set_option: bad argument");
    c->opt[option] = value;
    return MCL_OK;
}

extern "C" int64_t mcl_launch_count(mcl_ctx* c) { return c ? c->launches : 0; }

extern "C" int mcl_profile_enable(mcl_ctx* c, int on) {
    if (!c) return fail(MCL_EINVAL, "NULL context");
    c->profiling = on != 0;
    return MCL_OK;
}

extern "C" int mcl_profile_read(mcl_ctx* c, int which, double* ms, int64_t* launches) {
    if (!c || which < 0 || which >= PROF_N) return fail(MCL_EINVAL, "mcl_profile_read: bad argument");
    CU(cudaStreamSynchronize(c->stream));
    double total = 0;
    int64_t n = 0;
    for (auto& pr : c->prof[which]) {
        cudaEventSynchronize(pr.second);
        float t = 0;
        if (cudaEventElapsedTime(&t, pr.first, pr.second) == cudaSuccess) total += t;
        ++n;
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    c->prof[which].clear();
    if (ms) *ms = total;
    if (launches) *launches = n;
    return MCL_OK;
}

// ================================================================================================
// map index
// ================================================================================================
extern "C" void mcl_map_destroy(mcl_map* m) {
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    cudaFree(m->d_src_off);
    cudaFree(m->d_dst_off);
    cudaFree(m->d_desc);
    cudaFree(m->d_norm);
    cudaFree(m->d_img_of);
    for (auto& cs : m->chunk_sets) cudaFree(cs.d);
    cudaFree(m->d_meta3);
    cudaFree(m->d_tiles9);
    cudaFree(m->d_packed9);
    cudaFree(m->d_rec9);
    cudaFree(m->d_seg_rows);
    delete m;
}

extern "C" int mcl_map_create(mcl_ctx* c, int kind, int n_images, const int64_t* row_offsets, const void* desc,
                              int desc_dtype, mcl_map** out) {
    if (!c || !out) return fail(MCL_EINVAL, "mcl_map_create: NULL argument");
    *out = nullptr;
    if (kind != MCL_SIFT && kind != MCL_ORB && kind != MCL_SURF) return fail(MCL_EINVAL, "mcl_map_create: bad kind %d", kind);
    if (n_images < 0) return fail(MCL_EINVAL, "mcl_map_create: n_images < 0");
    if ((kind == MCL_ORB && desc_dtype != MCL_U8) || (kind == MCL_SURF && desc_dtype != MCL_F32) ||
        (desc_dtype != MCL_U8 && desc_dtype != MCL_F32))
        return fail(MCL_EINVAL, "mcl_map_create: dtype %d not valid for kind %d", desc_dtype, kind);
    int rc = check_offsets(row_offsets, n_images, "mcl_map_create");
    if (rc) return rc;
    const int64_t rows = row_offsets[n_images];
    if (rows > 0 && !desc) return fail(MCL_EINVAL, "mcl_map_create: desc is NULL");
    if (rows >= ((int64_t)1 << 31) - 4096) return fail(MCL_EINVAL, "mcl_map_create: too many rows for one shard");
    CU(cudaSetDevice(c->device));
    mcl_map* m = new (std::nothrow) mcl_map();
    if (!m) return fail(MCL_EINTERNAL, "out of host memory");
    m->ctx = c;
    m->kind = kind;
    m->n_images = n_images;
    m->rows = rows;
    m->h_src_off.assign(row_offsets, row_offsets + n_images + 1);
    for (int i = 0; i < n_images; ++i)
        m->max_rows_per_image = (int)std::max<int64_t>(m->max_rows_per_image, row_offsets[i + 1] - row_offsets[i]);
    cudaStream_t st = c->stream;
#define MCU(expr)                                                                                                 \
    do {                                                                                                          \
        cudaError_t e_ = (expr);                                                                                  \
        if (e_ != cudaSuccess) {                                                                                  \
            mcl_map_destroy(m);                                                                                   \
            return fail(MCL_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__);   \
        }                                                                                                         \
    } while (0)
    MCU(cudaMalloc(&m->d_src_off, (size_t)(n_images + 1) * 8));
    MCU(cudaMemcpyAsync(m->d_src_off, m->h_src_off.data(), (size_t)(n_images + 1) * 8, cudaMemcpyHostToDevice, st));
    if (kind == MCL_SIFT) {
        using namespace mcl::sifttc3;
        constexpr int TILE_ROWS = mcl::sifttc4::TILE_N;   // 128-row tiles
        m->h_dst_off.resize(n_images + 1);
        int64_t acc = 0;
        for (int i = 0; i < n_images; ++i) {
            m->h_dst_off[i] = acc;
            const int64_t r = row_offsets[i + 1] - row_offsets[i];
            acc += (r + GROUP - 1) / GROUP * GROUP;
        }
        m->h_dst_off[n_images] = acc;
        m->padded_rows = std::max<int64_t>(TILE_ROWS, (acc + TILE_ROWS - 1) / TILE_ROWS * TILE_ROWS);
        m->n_tiles = (int)(m->padded_rows / TILE_ROWS);
        m->n_units = (int)(m->padded_rows / TILE_N);
        MCU(cudaMalloc(&m->d_dst_off, (size_t)(n_images + 1) * 8));
        MCU(cudaMemcpyAsync(m->d_dst_off, m->h_dst_off.data(), (size_t)(n_images + 1) * 8, cudaMemcpyHostToDevice, st));
        MCU(cudaMalloc(&m->d_desc, (size_t)m->padded_rows * 128));
        MCU(cudaMalloc(&m->d_norm, (size_t)m->padded_rows * 4));
        MCU(cudaMalloc(&m->d_img_of, (size_t)m->padded_rows * 4));
        // rows of every image are stored sorted by |m|^2 (a 32-row group then has one norm, up to a few units)
        int32_t* d_perm = nullptr;
        if (rows > 0) {
            std::vector<int64_t> nrm((size_t)rows);
            if (desc_dtype == MCL_U8) {
                const uint8_t* p8 = (const uint8_t*)desc;
                for (int64_t r = 0; r < rows; ++r) {
                    int64_t a = 0;
                    for (int k = 0; k < 128; ++k) a += (int64_t)p8[r * 128 + k] * p8[r * 128 + k];
                    nrm[r] = a;
                }
            } else {
                const float* pf = (const float*)desc;
                for (int64_t r = 0; r < rows; ++r) {
                    double a = 0;
                    for (int k = 0; k < 128; ++k) a += (double)pf[r * 128 + k] * pf[r * 128 + k];
                    nrm[r] = (int64_t)a;
                }
            }
            std::vector<int32_t> perm((size_t)rows);
            for (int64_t r = 0; r < rows; ++r) perm[r] = (int32_t)r;
            for (int i = 0; i < n_images; ++i)
                std::stable_sort(perm.begin() + row_offsets[i], perm.begin() + row_offsets[i + 1],
                                 [&](int32_t a, int32_t b) { return nrm[a] < nrm[b]; });
            MCU(cudaMalloc(&d_perm, (size_t)rows * 4));
            MCU(cudaMemcpy(d_perm, perm.data(), (size_t)rows * 4, cudaMemcpyHostToDevice));
        }
        rc = ingest_rows128(c, st, desc, desc_dtype, rows, n_images, m->d_src_off, m->d_dst_off, m->padded_rows,
                            (uint8_t*)m->d_desc, m->d_norm, m->d_img_of, nullptr, true, d_perm);
        cudaStreamSynchronize(st);
        cudaFree(d_perm);
        if (rc) { mcl_map_destroy(m); return rc; }
        std::vector<Chunk> ch;
        int last_n = -1;
        // long chunks amortise sift_tc4's per-item pipeline refill; (448 + one image) x 2 records stay within its staged metadata
        for (int target : {448, 144, 64, 32, 16, 8, 4, 2, 1}) {
            build_chunks(m, TILE_ROWS, target, ch);
            if ((int)ch.size() == last_n) continue;
            last_n = (int)ch.size();
            mcl_map::ChunkSet cs;
            cs.target = target;
            cs.n = (int)ch.size();
            if (cs.n) {
                MCU(cudaMalloc(&cs.d, ch.size() * sizeof(Chunk)));
                m->chunk_sets.push_back(cs);   // owned by the map from here on (freed by mcl_map_destroy on a later failure)
                MCU(cudaMemcpy(cs.d, ch.data(), ch.size() * sizeof(Chunk), cudaMemcpyHostToDevice));
            } else {
                m->chunk_sets.push_back(cs);
            }
        }
        // per 64-row unit: image ids, descriptor counts and the groups' norm brackets
        MCU(cudaMalloc(&m->d_meta3, (size_t)m->n_units * sizeof(TileMeta)));
        {
            ProfScope ps(c, PROF_PREP, st);
            build_meta_kernel<<<(m->n_units + 127) / 128, 128, 0, st>>>(m->d_norm, m->d_img_of, m->d_src_off, m->n_units, m->d_meta3);
        }
        MCU(cudaGetLastError());
        m->resident_bytes = (size_t)m->padded_rows * (128 + 8) + (size_t)m->n_units * sizeof(TileMeta);
    } else {
        const size_t rb = row_bytes(kind, desc_dtype);
        m->padded_rows = rows;
        MCU(cudaMalloc(&m->d_desc, std::max<size_t>((size_t)rows * rb, 256)));
        if (rows) MCU(cudaMemcpyAsync(m->d_desc, desc, (size_t)rows * rb, cudaMemcpyHostToDevice, st));
        m->resident_bytes = (size_t)rows * rb;
        if (kind == MCL_ORB) {
            // tensor-core copy: every image padded to whole 32-row groups, the map to whole 128-row tiles of 256 signed bytes
            using namespace mcl::tc9;
            m->h_dst_off.resize(n_images + 1);
            int64_t acc = 0;
            std::vector<int32_t> seg_rows((size_t)std::max(1, n_images));
            for (int i = 0; i < n_images; ++i) {
                m->h_dst_off[i] = acc;
                const int64_t r = row_offsets[i + 1] - row_offsets[i];
                seg_rows[i] = (int32_t)r;
                acc += (r + GROUP - 1) / GROUP * GROUP;
            }
            m->h_dst_off[n_images] = acc;
            m->padded_rows = std::max<int64_t>(TILE_N, (acc + TILE_N - 1) / TILE_N * TILE_N);
            m->n_tiles = (int)(m->padded_rows / TILE_N);
            MCU(cudaMalloc(&m->d_dst_off, (size_t)(n_images + 1) * 8));
            MCU(cudaMemcpyAsync(m->d_dst_off, m->h_dst_off.data(), (size_t)(n_images + 1) * 8, cudaMemcpyHostToDevice, st));
            MCU(cudaMalloc(&m->d_seg_rows, seg_rows.size() * 4));
            MCU(cudaMemcpyAsync(m->d_seg_rows, seg_rows.data(), seg_rows.size() * 4, cudaMemcpyHostToDevice, st));
            MCU(cudaMalloc(&m->d_tiles9, (size_t)m->n_tiles * Shape<KIND_ORB>::TILE_BYTES));
            MCU(cudaMalloc(&m->d_packed9, (size_t)m->padded_rows * 32));
            MCU(cudaMalloc(&m->d_img_of, (size_t)m->padded_rows * 4));
            MCU(cudaMalloc(&m->d_rec9, (size_t)m->n_tiles * sizeof(TileRec)));
            MCU(cudaMemsetAsync(m->d_tiles9, 0, (size_t)m->n_tiles * Shape<KIND_ORB>::TILE_BYTES, st));
            MCU(cudaMemsetAsync(m->d_packed9, 0, (size_t)m->padded_rows * 32, st));
            MCU(cudaMemsetAsync(m->d_img_of, 0xFF, (size_t)m->padded_rows * 4, st));
            {
                ProfScope ps(c, PROF_PREP, st);
                if (rows) orb_map_prep_kernel<<<grid_for(rows * 32, 256, c->sm_count * 8), 256, 0, st>>>(
                    (const uint8_t*)m->d_desc, rows, m->d_src_off, m->d_dst_off, n_images, m->d_tiles9, m->d_packed9, m->d_img_of);
                orb_build_rec_kernel<<<(m->n_tiles + 127) / 128, 128, 0, st>>>(m->d_img_of, m->d_src_off, m->d_dst_off, m->n_tiles, m->d_rec9);
            }
            MCU(cudaGetLastError());
            std::vector<mcl::sifttc3::Chunk> ch;
            int last_n = -1;
            for (int target : {512, 128, 32, 8, 4, 2, 1}) {
                build_chunks(m, TILE_N, target, ch);
                if ((int)ch.size() == last_n) continue;
                last_n = (int)ch.size();
                mcl_map::ChunkSet cs;
                cs.target = target;
                cs.n = (int)ch.size();
                if (cs.n) {
                    MCU(cudaMalloc(&cs.d, ch.size() * sizeof(mcl::sifttc3::Chunk)));
                    m->chunk_sets.push_back(cs);
                    MCU(cudaMemcpy(cs.d, ch.data(), ch.size() * sizeof(mcl::sifttc3::Chunk), cudaMemcpyHostToDevice));
                } else {
                    m->chunk_sets.push_back(cs);
                }
            }
            m->resident_bytes += (size_t)m->n_tiles * Shape<KIND_ORB>::TILE_BYTES + (size_t)m->padded_rows * 36;
        }
    }
    MCU(cudaStreamSynchronize(st));
#undef MCU
    *out = m;
    return MCL_OK;
}

extern "C" int mcl_map_info(mcl_map* m, int64_t info[4]) {
    if (!m || !info) return fail(MCL_EINVAL, "mcl_map_info: NULL argument");
    info[0] = m->n_images;
    info[1] = m->rows;
    info[2] = m->padded_rows;
    info[3] = (int64_t)m->resident_bytes;
    return MCL_OK;
}

// ================================================================================================
// query batch
// ================================================================================================
extern "C" void mcl_queries_destroy(mcl_queries* q) {
    if (!q) return;
    if (q->pooled) {
        cudaSetDevice(q->ctx->device);
        cudaStreamSynchronize(q->built_on);
        cudaStreamSynchronize(q->ctx->stream);
        pool_free(q->ctx, q->d_off, q->sz_off);
        pool_free(q->ctx, q->d_desc, q->sz_desc);
        pool_free(q->ctx, q->d_norm, q->sz_norm);
        pool_free(q->ctx, q->d_frame, q->sz_frame);
        pool_free(q->ctx, q->d_bad, 256);
    } else if (!q->borrowed) {
        cudaSetDevice(q->ctx->device);
        cudaStreamSynchronize(q->ctx->stream);
        cudaFree(q->d_off);
        cudaFree(q->d_desc);
        cudaFree(q->d_norm);
        cudaFree(q->d_frame);
        cudaFree(q->d_bad);
    }
    delete q;
}

namespace {
// borrowed = true: device buffers come from the context's growable scratch (the per-call batch of mcl_match_counts /
// mcl_bow_score: no cudaMalloc / cudaFree on the per-frame path); false: owned allocations (resident batch).
int queries_build(mcl_ctx* c, int kind, int n_queries, const int64_t* q_row_offsets, const void* q_desc, int q_dtype,
                  bool borrowed, mcl_queries** out, bool src_on_device = false, cudaStream_t on_stream = nullptr) {
    *out = nullptr;
    if (kind < MCL_SIFT || kind > MCL_BOWQ) return fail(MCL_EINVAL, "queries: bad kind %d", kind);
    if (n_queries < 0) return fail(MCL_EINVAL, "queries: n_queries < 0");
    if ((kind == MCL_ORB && q_dtype != MCL_U8) || (kind == MCL_SURF && q_dtype != MCL_F32) ||
        (q_dtype != MCL_U8 && q_dtype != MCL_F32))
        return fail(MCL_EINVAL, "queries: dtype %d not valid for kind %d", q_dtype, kind);
    int rc = check_offsets(q_row_offsets, n_queries, "queries");
    if (rc) return rc;
    const int64_t rows = q_row_offsets[n_queries];
    if (rows > 0 && !q_desc) return fail(MCL_EINVAL, "queries: q_desc is NULL");
    if (rows >= ((int64_t)1 << 31) - 4096) return fail(MCL_EINVAL, "queries: too many rows");
    CU(cudaSetDevice(c->device));
    mcl_queries* q = new (std::nothrow) mcl_queries();
    if (!q) return fail(MCL_EINTERNAL, "out of host memory");
    q->ctx = c;
    q->kind = kind;
    q->n_queries = n_queries;
    q->rows = rows;
    q->borrowed = borrowed;
    q->h_off.assign(q_row_offsets, q_row_offsets + n_queries + 1);
    cudaStream_t st = src_on_device && on_stream ? on_stream : c->stream;
    q->built_on = st;
    const cudaMemcpyKind h2d = src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
#define QCU(expr)                                                                                                 \
    do {                                                                                                          \
        cudaError_t e_ = (expr);                                                                                  \
        if (e_ != cudaSuccess) {                                                                                  \
            mcl_queries_destroy(q);                                                                               \
            return fail(MCL_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__);   \
        }                                                                                                         \
    } while (0)
    q->pooled = src_on_device;
    auto get = [&](DevBuf& scratch, void** dst, size_t bytes) -> cudaError_t {
        bytes = std::max<size_t>(bytes, 256);
        if (borrowed) {
            cudaError_t e = scratch.ensure(bytes);
            *dst = scratch.p;
            return e;
        }
        if (q->pooled) {
            (dst == (void**)&q->d_off ? q->sz_off : dst == &q->d_desc ? q->sz_desc : dst == (void**)&q->d_norm ? q->sz_norm : q->sz_frame) = bytes;
            return pool_alloc(c, dst, bytes);
        }
        return cudaMalloc(dst, bytes);
    };
    QCU(get(c->tq_off, (void**)&q->d_off, (size_t)(n_queries + 1) * 8));
    QCU(cudaMemcpyAsync(q->d_off, q->h_off.data(), (size_t)(n_queries + 1) * 8, cudaMemcpyHostToDevice, st));
    if (kind == MCL_SIFT) {
        q->padded_rows = std::max<int64_t>(QROW_PAD, (rows + QROW_PAD - 1) / QROW_PAD * QROW_PAD);
        QCU(get(c->tq_desc, &q->d_desc, (size_t)q->padded_rows * 128));
        QCU(get(c->tq_norm, (void**)&q->d_norm, (size_t)q->padded_rows * 4));
        QCU(get(c->tq_frame, (void**)&q->d_frame, (size_t)q->padded_rows * 4));
        if (src_on_device) {
            QCU(pool_alloc(c, (void**)&q->d_bad, 256));
            QCU(cudaMemsetAsync(q->d_bad, 0, 4, st));
        }
        rc = ingest_rows128(c, st, q_desc, q_dtype, rows, n_queries, q->d_off, q->d_off, q->padded_rows,
                            (uint8_t*)q->d_desc, q->d_norm, q->d_frame, nullptr, !src_on_device, nullptr, nullptr, nullptr,
                            src_on_device, q->d_bad);
        if (rc) { mcl_queries_destroy(q); return rc; }
    } else if (kind == MCL_BOWQ) {
        q->padded_rows = rows;
        QCU(get(c->tq_desc, &q->d_desc, (size_t)rows * 512));
        if (rows) {
            if (q_dtype == MCL_F32) {
                QCU(cudaMemcpyAsync(q->d_desc, q_desc, (size_t)rows * 512, h2d, st));
            } else {
                QCU(c->stage.ensure((size_t)rows * 128));
                QCU(cudaMemcpyAsync(c->stage.p, q_desc, (size_t)rows * 128, h2d, st));
                ProfScope ps(c, PROF_PREP, st);
                mcl::u8_to_f32_kernel<<<grid_for(rows * 128, 256, c->sm_count * 8), 256, 0, st>>>(
                    c->stage.as<uint8_t>(), (float*)q->d_desc, rows * 128);
            }
        }
    } else {
        const size_t rb = row_bytes(kind, q_dtype);
        q->padded_rows = rows;
        QCU(get(c->tq_desc, &q->d_desc, (size_t)rows * rb));
        if (rows) QCU(cudaMemcpyAsync(q->d_desc, q_desc, (size_t)rows * rb, h2d, st));
    }
    // the host offsets/descriptors are caller-owned: the copies must have been consumed before we return (a device-built
    // batch only enqueues: its offsets live in q->h_off, which outlives the copy)
    if (!src_on_device) QCU(cudaStreamSynchronize(st));
#undef QCU
    *out = q;
    return MCL_OK;
}
}  // namespace

extern "C" int mcl_queries_create(mcl_ctx* c, int kind, int n_queries, const int64_t* q_row_offsets, const void* q_desc,
                                  int q_dtype, mcl_queries** out) {
    if (!c || !out) return fail(MCL_EINVAL, "mcl_queries_create: NULL argument");
    return queries_build(c, kind, n_queries, q_row_offsets, q_desc, q_dtype, false, out);
}

extern "C" int mcl_queries_create_device(mcl_ctx* c, int kind, int n_queries, const int64_t* q_row_offsets, const void* d_desc,
                                         int q_dtype, void* stream, mcl_queries** out) {
    if (!c || !out) return fail(MCL_EINVAL, "mcl_queries_create_device: NULL argument");
    return queries_build(c, kind, n_queries, q_row_offsets, d_desc, q_dtype, false, out, true, (cudaStream_t)stream);
}

extern "C" int mcl_queries_validate(mcl_queries* q) {
    if (!q) return fail(MCL_EINVAL, "mcl_queries_validate: NULL argument");
    if (!q->d_bad) return MCL_OK;   // host-built batches were checked when they were created
    int h_bad = 0;
    CU(cudaSetDevice(q->ctx->device));
    CU(cudaMemcpyAsync(&h_bad, q->d_bad, 4, cudaMemcpyDeviceToHost, q->built_on));
    CU(cudaStreamSynchronize(q->built_on));
    if (h_bad) return fail(MCL_EDATA, "SIFT descriptors must be integer valued in [0,255]");
    return MCL_OK;
}

// ================================================================================================
// matching
// ================================================================================================
namespace {

// cluster launch helper
template <typename K, typename P>
cudaError_t launch_cluster(K kernel, int grid, int threads, size_t smem, cudaStream_t st, int cl, const P& p) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cl;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, p);
}
template <typename K>
int max_active_clusters(K kernel, int sm_count, int threads, size_t smem, int cl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((sm_count / cl) * cl);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cl;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) { n = 0; cudaGetLastError(); }
    return n;
}

// one-time attributes + co-residency of the tc9 kernels
int tc9_prepare(mcl_ctx* c) {
    using namespace mcl::tc9;
    if (c->tc9_attr) return MCL_OK;
    CU(cudaFuncSetAttribute(tc9_kernel<KIND_ORB, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<KIND_ORB>)));
    CU(cudaFuncSetAttribute(tc9_kernel<KIND_ORB, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<KIND_ORB>)));
    CU(cudaFuncSetAttribute(tc9_kernel<KIND_BOW, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<KIND_BOW>)));
    CU(cudaFuncSetAttribute(tc9_kernel<KIND_BOW, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<KIND_BOW>)));
    c->tc9_clusters[0] = max_active_clusters(tc9_kernel<KIND_ORB, 2>, c->sm_count, THREADS, sizeof(Smem<KIND_ORB>), 2);
    c->tc9_clusters[1] = max_active_clusters(tc9_kernel<KIND_BOW, 2>, c->sm_count, THREADS, sizeof(Smem<KIND_BOW>), 2);
    if (getenv("MCL_DEBUG")) fprintf(stderr, "[mcl] co-resident tc9 CTA pairs: ORB %d, BOW %d\n", c->tc9_clusters[0], c->tc9_clusters[1]);
    c->tc9_attr = true;
    return MCL_OK;
}

// MCL_TC9_DIAG: where the MMA warp of tc9_kernel waited (diagnostic; synchronises the stream)
int tc9_diag_report(mcl_ctx* c, cudaStream_t st, const char* what, const long long* d_diag, int grid, int n_chunks, int n_qgroups) {
    std::vector<long long> h((size_t)grid * 5);
    CU(cudaStreamSynchronize(st));
    CU(cudaMemcpy(h.data(), d_diag, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    double a = 0, b = 0, ac = 0, tot = 0, tiles = 0, tmax = 0;
    for (int i = 0; i < grid; ++i) { a += h[5 * i]; b += h[5 * i + 1]; ac += h[5 * i + 2]; tot += h[5 * i + 3]; tiles += h[5 * i + 4]; tmax = std::max(tmax, (double)h[5 * i + 3]); }
    fprintf(stderr, "[mcl] tc9 %s: %d CTAs, %d chunks x %d query groups; per CTA: %.0f tiles, %.0f cycles (max %.0f) = %.0f per tile; "
                    "MMA warp waits: query tiles %.1f %%, map tiles %.1f %%, accumulators %.1f %%\n",
            what, grid, n_chunks, n_qgroups, tiles / grid, tot / grid, tmax, tot / std::max(1.0, tiles), 100 * a / tot, 100 * b / tot, 100 * ac / tot);
    (void)c;
    return MCL_OK;
}

// SIFT on the tensor cores: sift_tc4_kernel (bracketed pre-filter, candidates) + sift_fixup3_kernel (exact resolution) per
// batch of query blocks, then the exact SIMT kernel on the (frame, image) pairs that lost a candidate to a full list.
int launch_sift_tc4(mcl_map* m, mcl_queries* q, double ratio, int32_t* d_counts, cudaStream_t st) {
    using mcl::sifttc3::Params;
    using mcl::sifttc3::sift_fixup3_kernel;
    using namespace mcl::sifttc4;
    mcl_ctx* c = m->ctx;
    constexpr size_t smem = sizeof(SmemT<0>);
    if (!c->tc4_attr) {
        CU(cudaFuncSetAttribute(sift_tc4_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU(cudaFuncSetAttribute(sift_tc4_kernel<0, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU(cudaFuncSetAttribute(sift_tc4_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU(cudaFuncSetAttribute(sift_tc4_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU(cudaFuncSetAttribute(sift_tc4_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // how many CTA pairs are co-resident (GPC boundaries may strand SMs)
        c->max_clusters2 = max_active_clusters(sift_tc4_kernel<0, 0, 2>, c->sm_count, mcl::sifttc4::THREADS, smem, 2);
        if (getenv("MCL_DEBUG")) fprintf(stderr, "[mcl] co-resident sift_tc4 clusters of 2 CTAs: %d (%d SMs)\n", c->max_clusters2, c->sm_count);
        c->tc4_attr = true;
    }
    const int n_qblocks_total = (int)(q->padded_rows / QBLK);
    // candidate list: sized for 1/4 of all (query row, image) pairs of a batch being candidates.  A candidate that does not fit
    // marks its (frame, image) pair dirty; dirty pairs are recomputed by the exact SIMT kernel at the end of the call, so
    // the size is only a performance knob (an adversarial batch degrades pair by pair, not as a whole).
    const int64_t pairs_total = q->padded_rows * (int64_t)std::max(1, m->n_images);
    const int64_t want = std::min<int64_t>(std::max<int64_t>(pairs_total / 4, 1 << 16), (int64_t)48 << 20);
    if ((size_t)want > c->cand_cap_entries) {
        CU(c->cand.ensure((size_t)want * sizeof(int4)));
        c->cand_cap_entries = (size_t)want;
    }
    // option 3 (tests): cap the candidate list to force the dirty-pair repair path
    const int64_t cap = c->opt[3] > 0 ? std::min<int64_t>(c->opt[3], (int64_t)c->cand_cap_entries) : (int64_t)c->cand_cap_entries;
    int batch_qblocks = (int)std::max<int64_t>(1, cap * 4 / ((int64_t)std::max(1, m->n_images) * QBLK));
    batch_qblocks = std::min(batch_qblocks, n_qblocks_total);
    unsigned int* d_cand_count = c->flags.as<unsigned int>();
    int* d_overflow = c->flags.as<int>() + 1;
    unsigned int* d_rescans = c->flags.as<unsigned int>() + 3;
    unsigned long long* d_stats = c->flags.as<unsigned long long>() + 8;   // [0] candidates, [1] dirty pairs (bytes 64..79)
    const size_t n_pairs = (size_t)std::max(1, q->n_queries) * std::max(1, m->n_images);
    unsigned int* d_nlist = c->flags.as<unsigned int>() + 4;   // entries of the dirty-pair list
    CU(c->dirty.ensure(n_pairs));
    CU(c->dirty_list.ensure(n_pairs * sizeof(int)));
    CU(cudaMemsetAsync(c->dirty.p, 0, n_pairs, st));
    CU(cudaMemsetAsync(d_overflow, 0, 4, st));
    CU(cudaMemsetAsync(d_nlist, 0, 4, st));
    for (int qb0 = 0; qb0 < n_qblocks_total; qb0 += batch_qblocks) {
        const int nqb = std::min(batch_qblocks, n_qblocks_total - qb0);
        const mcl_map::ChunkSet* cs = pick_chunks(m->chunk_sets, nqb, c->sm_count, m->n_tiles);
        if (cs->n == 0) break;
        CU(cudaMemsetAsync(d_cand_count, 0, 4, st));
        Params p;
        p.q_desc = (const uint8_t*)q->d_desc + (size_t)qb0 * QBLK * 128;
        p.q_norm = q->d_norm + (size_t)qb0 * QBLK;
        p.q_frame = q->d_frame + (size_t)qb0 * QBLK;
        p.m_desc = (const uint8_t*)m->d_desc;
        p.m_meta = m->d_meta3;
        p.chunks = cs->d;
        p.n_qblocks = nqb;
        p.n_chunks = cs->n;
        p.ratio = ratio;
        p.cand = c->cand.as<int4>();
        p.cand_count = d_cand_count;
        p.cand_cap = (unsigned)cap;
        p.overflow = d_overflow;
        p.dirty = c->dirty.as<uint8_t>();
        p.n_images = m->n_images;
        p.q_row_base = qb0 * QBLK;
        p.meta_cap = c->opt[1] == 1 ? 0 : META_CAP;   // option 1 = 1: always read the tile records from global memory
        // CTAs of a cluster share the map tiles (multicast): less L2 traffic -> less power; a query-block count that is not
        // a multiple of the cluster size idles one CTA for one item per chunk, so small odd batches launch plain CTAs.
        // option 5: 0 = automatic, 1 = never, 2 = always pairs
        int cl = 1;
        if (c->opt[0] == 0) {
            if (c->opt[5] == 2) cl = 2;
            else if (c->opt[5] == 0 && ((nqb % 2 == 0 && nqb >= 2) || nqb >= 32)) cl = 2;
            if (cl > 1 && c->max_clusters2 < 1) cl = 1;
        }
        const int64_t n_items = (int64_t)((nqb + cl - 1) / cl) * cs->n;
        const int grid = cl > 1 ? cl * (int)std::min<int64_t>(n_items, c->max_clusters2) : (int)std::min<int64_t>(n_items, c->sm_count);
        p.diag = nullptr;
        #ifdef MCL_TC_DIAG
        static const bool diag4 = getenv("MCL_TC9_DIAG") != nullptr;
#else
        constexpr bool diag4 = false;
#endif
        if (diag4 && c->opt[0] == 0) {
            CU(c->diag9.ensure((size_t)grid * 5 * sizeof(long long)));
            CU(cudaMemsetAsync(c->diag9.p, 0, (size_t)grid * 5 * sizeof(long long), st));
            p.diag = c->diag9.as<long long>();
        }
        {
            ProfScope ps(c, PROF_SIFT_TC, st);
            const int T = mcl::sifttc4::THREADS;
            if (c->opt[0] == 3) sift_tc4_kernel<3><<<grid, T, smem, st>>>(p);
            else if (c->opt[0] == 4) sift_tc4_kernel<4><<<grid, T, smem, st>>>(p);
            else if (c->opt[0] == 5) {   // timeline of 64 consecutive tiles: epilogue warp 5 and the MMA warp of CTA 0
                sift_tc4_kernel<5><<<grid, T, smem, st>>>(p);
                long long tl[64 * 5], tm[64 * 6];
                cudaStreamSynchronize(st);
                cudaMemcpy(tl, reinterpret_cast<long long*>(c->cand.as<int4>() + (cap - 1024)), sizeof tl, cudaMemcpyDeviceToHost);
                cudaMemcpy(tm, reinterpret_cast<long long*>(c->cand.as<int4>() + (cap - 1024)) + 1024, sizeof tm, cudaMemcpyDeviceToHost);
                for (int i = 0; i < 64; ++i)
                    fprintf(stderr, "tile %2d epi: wait_full %5lld ld1 %5lld max+ld2+arrive %5lld compute %5lld period %5lld | mma: wait_b %5lld wait_e0 %5lld e1 %5lld e2 %5lld rest %5lld period %5lld | full->empty %5lld\n", i,
                            tl[i * 5 + 1] - tl[i * 5], tl[i * 5 + 2] - tl[i * 5 + 1], tl[i * 5 + 3] - tl[i * 5 + 2], tl[i * 5 + 4] - tl[i * 5 + 3],
                            i ? tl[i * 5] - tl[i * 5 - 5] : 0, tm[i * 6 + 1] - tm[i * 6], tm[i * 6 + 2] - tm[i * 6 + 1], tm[i * 6 + 3] - tm[i * 6 + 2],
                            tm[i * 6 + 4] - tm[i * 6 + 3], tm[i * 6 + 5] - tm[i * 6 + 4], i ? tm[i * 6] - tm[i * 6 - 6] : 0,
                            tl[i * 5 + 3] - tl[i * 5 + 1]);
            }
            else if (cl == 2) CU(launch_cluster(sift_tc4_kernel<0, 0, 2>, grid, T, smem, st, 2, p));
            else sift_tc4_kernel<0><<<grid, T, smem, st>>>(p);
        }
        CU(cudaGetLastError());
        if (p.diag) {
            const int r = tc9_diag_report(c, st, "SIFT (sift_tc4, tiles of 64 rows x 3 query tiles)", p.diag, grid, cs->n, (int)((nqb + cl - 1) / cl));
            if (r) return r;
        }
        {
            ProfScope ps(c, PROF_SIFT_FIX, st);
            // grid-stride over the candidates: exactly the co-resident blocks, so that no block starts a wave late
            auto fix = sift_fixup3_kernel;
            if (!c->fixup_blocks) {
                int per_sm = 0;
                CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fix, 256, 0));
                c->fixup_blocks = c->sm_count * std::max(1, per_sm);
            }
            fix<<<c->fixup_blocks, 256, 0, st>>>(c->cand.as<int4>(), d_cand_count, (unsigned)cap, (const uint8_t*)q->d_desc, q->d_norm,
                                                 q->d_frame, (const uint8_t*)m->d_desc, m->d_norm, m->d_img_of, m->d_dst_off, ratio,
                                                 m->n_images, d_counts, d_rescans, d_stats);
        }
        CU(cudaGetLastError());
    }
    // exact repair of the dirty (frame, image) pairs, if any: the SIMT kernel overwrites their counts
    {
        ProfScope ps(c, PROF_SIFT_SIMT, st);
        mcl::compact_dirty_kernel<<<grid_for((int64_t)n_pairs, 256, c->sm_count * 4), 256, 0, st>>>(
            c->dirty.as<uint8_t>(), (int64_t)n_pairs, d_overflow, c->dirty_list.as<int>(), d_nlist);
        c->launches++;   // two kernels in this scope
        mcl::sift_simt_list_kernel<<<c->sm_count * 8, 128, 0, st>>>(c->dirty_list.as<int>(), d_nlist, (const uint8_t*)q->d_desc, q->d_norm,
                                                                   q->d_off, (const uint8_t*)m->d_desc, m->d_norm, m->d_dst_off,
                                                                   m->d_src_off, m->n_images, ratio, d_counts, d_stats + 1);
    }
    CU(cudaGetLastError());
    return MCL_OK;
}

// ORB 2-NN + ratio on the tensor cores: tc9_kernel<KIND_ORB> (exact best key, lower bound of the second) + orb_fixup_kernel
// (xor + popc on the best group) per batch of query blocks, then the SIMT kernel on the (frame, image) pairs that lost a
// candidate to a full list.
int launch_orb_tc(mcl_map* m, mcl_queries* q, double ratio, int32_t* d_counts, cudaStream_t st) {
    using namespace mcl::tc9;
    mcl_ctx* c = m->ctx;
    int rc = tc9_prepare(c);
    if (rc) return rc;
    const int n_qblocks_total = (int)((q->rows + QBLK - 1) / QBLK);
    const int64_t padded = (int64_t)n_qblocks_total * QBLK;
    CU(c->orb_q.ensure((size_t)padded * (256 + 8)));
    uint8_t* q_rows = c->orb_q.as<uint8_t>();
    int32_t* q_aux = reinterpret_cast<int32_t*>(q_rows + (size_t)padded * 256);
    int32_t* q_frame = q_aux + padded;
    CU(cudaMemsetAsync(q_rows, 0, (size_t)padded * 256, st));
    CU(cudaMemsetAsync(q_aux, 0xFF, (size_t)padded * 8, st));
    {
        ProfScope ps(c, PROF_PREP, st);
        orb_query_prep_kernel<<<grid_for(q->rows * 32, 256, c->sm_count * 8), 256, 0, st>>>((const uint8_t*)q->d_desc, q->rows, q->d_off,
                                                                                         q->n_queries, q_rows, q_aux, q_frame);
    }
    CU(cudaGetLastError());
    const int64_t pairs_total = padded * (int64_t)std::max(1, m->n_images);
    const int64_t want = std::min<int64_t>(std::max<int64_t>(pairs_total / 4, 1 << 16), (int64_t)48 << 20);
    if ((size_t)want > c->cand_cap_entries) {
        CU(c->cand.ensure((size_t)want * sizeof(int4)));
        c->cand_cap_entries = (size_t)want;
    }
    const int64_t cap = c->opt[3] > 0 ? std::min<int64_t>(c->opt[3], (int64_t)c->cand_cap_entries) : (int64_t)c->cand_cap_entries;
    int batch_qblocks = (int)std::max<int64_t>(1, cap * 4 / ((int64_t)std::max(1, m->n_images) * QBLK));
    batch_qblocks = std::min(batch_qblocks, n_qblocks_total);
    unsigned int* d_cand_count = c->flags.as<unsigned int>();
    int* d_overflow = c->flags.as<int>() + 1;
    unsigned long long* d_stats = c->flags.as<unsigned long long>() + 8;
    const size_t n_pairs = (size_t)std::max(1, q->n_queries) * std::max(1, m->n_images);
    unsigned int* d_nlist = c->flags.as<unsigned int>() + 4;   // entries of the dirty-pair list
    CU(c->dirty.ensure(n_pairs));
    CU(c->dirty_list.ensure(n_pairs * sizeof(int)));
    CU(cudaMemsetAsync(c->dirty.p, 0, n_pairs, st));
    CU(cudaMemsetAsync(d_overflow, 0, 4, st));
    CU(cudaMemsetAsync(d_nlist, 0, 4, st));
    for (int qb0 = 0; qb0 < n_qblocks_total; qb0 += batch_qblocks) {
        const int nqb = std::min(batch_qblocks, n_qblocks_total - qb0);
        const mcl_map::ChunkSet* cs = pick_chunks(m->chunk_sets, nqb, c->sm_count, m->n_tiles, 1.5);
        if (cs->n == 0) break;
        CU(cudaMemsetAsync(d_cand_count, 0, 4, st));
        Params p = {};
        p.q_rows = q_rows + (size_t)qb0 * QBLK * 256;
        p.q_aux = q_aux + (size_t)qb0 * QBLK;
        p.q_frame = q_frame + (size_t)qb0 * QBLK;
        p.m_tiles = m->d_tiles9;
        p.m_rec = m->d_rec9;
        p.chunks = cs->d;
        p.seg_rows = m->d_seg_rows;
        p.n_qblocks = nqb;
        p.n_chunks = cs->n;
        p.ratio = ratio;
        p.cand = c->cand.as<int4>();
        p.cand_count = d_cand_count;
        p.cand_cap = (unsigned)cap;
        p.overflow = d_overflow;
        p.dirty = c->dirty.as<uint8_t>();
        p.n_images = m->n_images;
        p.q_row_base = qb0 * QBLK;
        p.meta_cap = c->opt[1] == 1 ? 0 : META_CAP;
        int cl = 1;
        if (c->opt[5] == 2 || (c->opt[5] == 0 && ((nqb % 2 == 0 && nqb >= 2) || nqb >= 32))) cl = 2;
        if (cl == 2 && c->tc9_clusters[0] < 1) cl = 1;
        const int64_t n_items = (int64_t)((nqb + cl - 1) / cl) * cs->n;
        const int grid9 = cl == 2 ? 2 * (int)std::min<int64_t>(n_items, c->tc9_clusters[0]) : (int)std::min<int64_t>(n_items, c->sm_count);
        #ifdef MCL_TC_DIAG
        static const bool diag9 = getenv("MCL_TC9_DIAG") != nullptr;
#else
        constexpr bool diag9 = false;
#endif
        if (diag9) {
            CU(c->diag9.ensure((size_t)grid9 * 5 * sizeof(long long)));
            CU(cudaMemsetAsync(c->diag9.p, 0, (size_t)grid9 * 5 * sizeof(long long), st));
            p.diag = c->diag9.as<long long>();
        }
        {
            ProfScope ps(c, PROF_ORB, st);
            constexpr size_t smem = sizeof(Smem<KIND_ORB>);
            if (cl == 2) CU(launch_cluster(tc9_kernel<KIND_ORB, 2>, grid9, THREADS, smem, st, 2, p));
            else tc9_kernel<KIND_ORB, 1><<<grid9, THREADS, smem, st>>>(p);
        }
        if (diag9) {
            const int r = tc9_diag_report(c, st, "ORB", p.diag, grid9, cs->n, (int)((nqb + cl - 1) / cl));
            if (r) return r;
        }
        CU(cudaGetLastError());
        {
            ProfScope ps(c, PROF_ORB_FIX, st);
            orb_fixup_kernel<<<c->sm_count * 4, 256, 0, st>>>(c->cand.as<int4>(), d_cand_count, (unsigned)cap, (const uint4*)q->d_desc,
                                                             q_frame, (const uint4*)m->d_packed9, m->d_src_off, m->d_dst_off, ratio,
                                                             m->n_images, d_counts, d_stats + 2);
        }
        CU(cudaGetLastError());
    }
    {   // exact repair of the dirty (frame, image) pairs, if any: their counts are reset and recomputed by the SIMT kernel
        ProfScope ps(c, PROF_ORB_FIX, st);
        zero_where_mask_kernel<<<grid_for((int64_t)n_pairs, 256, c->sm_count * 4), 256, 0, st>>>(d_counts, c->dirty.as<uint8_t>(), (int64_t)n_pairs, d_overflow);
        int64_t max_q = 0;
        for (int f = 0; f < q->n_queries; ++f) max_q = std::max(max_q, q->h_off[f + 1] - q->h_off[f]);
        mcl::compact_dirty_kernel<<<grid_for((int64_t)n_pairs, 256, c->sm_count * 4), 256, 0, st>>>(
            c->dirty.as<uint8_t>(), (int64_t)n_pairs, d_overflow, c->dirty_list.as<int>(), d_nlist);
        c->launches += 2;   // three kernels in this scope
        mcl::orb_knn_list_kernel<<<c->sm_count * 8, 128, 0, st>>>(c->dirty_list.as<int>(), d_nlist, (int)std::max<int64_t>(1, (max_q + 127) / 128),
                                                                 (const uint4*)q->d_desc, q->d_off, (const uint4*)m->d_desc, m->d_src_off,
                                                                 m->n_images, ratio, d_counts, d_stats + 1);
    }
    c->launches++;
    CU(cudaGetLastError());
    return MCL_OK;
}

int match_device(mcl_map* m, mcl_queries* q, int mode, double ratio, const uint8_t* d_mask, int32_t fill, int algo,
                 int32_t* d_counts, cudaStream_t st) {
    mcl_ctx* c = m->ctx;
    const int nf = q->n_queries, ni = m->n_images;
    const int64_t n_out = (int64_t)nf * ni;
    if (n_out == 0) return MCL_OK;
    if (nf > 65535) return fail(MCL_EINVAL, "at most 65535 query frames per call (got %d)", nf);
    CU(cudaMemsetAsync(d_counts, 0, (size_t)n_out * 4, st));
    if (q->rows > 0 && m->rows > 0) {
        if (m->kind == MCL_SIFT) {
            if (mode != MCL_KNN2_RATIO) return fail(MCL_EINVAL, "SIFT supports mode MCL_KNN2_RATIO only");
            int use = algo;
            if (use == MCL_ALGO_AUTO) {
                // measured (scripts/latency_probe.py): the tensor-core path wins even for one 150-row frame against 25 small
                // images (93 us vs 122 us).  Only DOR-masked calls on small problems go to the SIMT kernel, which skips the
                // masked pairs (<= 30 of 175 images per frame) instead of matching everything and masking afterwards.
                const double pairs = (double)q->rows * (double)m->rows;
                use = (d_mask && pairs < 2.0e7) ? MCL_ALGO_SIMT : MCL_ALGO_TENSOR;
            }
            if (use == MCL_ALGO_TENSOR) {
                int rc = launch_sift_tc4(m, q, ratio, d_counts, st);
                if (rc) return rc;
            } else {
                ProfScope ps(c, PROF_SIFT_SIMT, st);
                dim3 grid(ni, nf);
                mcl::sift_simt_kernel<<<grid, 128, 0, st>>>((const uint8_t*)q->d_desc, q->d_norm, q->d_off,
                                                            (const uint8_t*)m->d_desc, m->d_norm, m->d_dst_off, m->d_src_off,
                                                            ni, ratio, d_mask, d_counts, nullptr);
            }
            CU(cudaGetLastError());
        } else {
            int64_t max_q = 0;
            for (int f = 0; f < nf; ++f) max_q = std::max(max_q, q->h_off[f + 1] - q->h_off[f]);
            // thread per query row; a small SURF problem (one frame against a DOR window) gets narrower blocks so that the
            // blocks spread evenly over the SMs (0.46 -> 0.39 ms for 1 x 30 x 1024 rows).  Not ORB: every block stages the
            // whole map image in shared memory, and for 32-byte rows that fixed cost outweighs the better balance (measured)
            int bt = 128;
            while (m->kind == MCL_SURF && bt > 32 && (int64_t)ni * nf * ((max_q + bt - 1) / bt) < (int64_t)4 * c->sm_count) bt >>= 1;
            const int zq = (int)std::max<int64_t>(1, (max_q + bt - 1) / bt);
            if (zq > 65535 || ni > 2147483647) return fail(MCL_EINVAL, "query frame too large");
            dim3 grid(ni, nf, zq);
            if (m->kind == MCL_ORB) {
                // 2-NN + ratio: the tensor cores unless the call is small or DOR-masked (the SIMT kernel skips masked pairs)
                const bool tensor = mode == MCL_KNN2_RATIO &&
                                    (algo == MCL_ALGO_TENSOR || (algo == MCL_ALGO_AUTO && !d_mask && (double)q->rows * (double)m->rows >= 2.0e6));
                if (tensor) {
                    int rc = launch_orb_tc(m, q, ratio, d_counts, st);
                    if (rc) return rc;
                } else if (mode == MCL_KNN2_RATIO) {
                    ProfScope ps(c, PROF_ORB, st);
                    mcl::orb_knn_kernel<<<grid, bt, 0, st>>>((const uint4*)q->d_desc, q->d_off, (const uint4*)m->d_desc,
                                                              m->d_src_off, ni, ratio, d_mask, d_counts, nullptr);
                } else if (mode == MCL_CROSSCHECK) {
                    CU(c->row_arg.ensure((size_t)q->rows * ni * 4));
                    {
                        ProfScope ps(c, PROF_ORB, st);
                        mcl::orb_rowargmin_kernel<<<grid, bt, 0, st>>>((const uint4*)q->d_desc, q->d_off,
                                                                        (const uint4*)m->d_desc, m->d_src_off, ni, d_mask,
                                                                        c->row_arg.as<int32_t>());
                    }
                    CU(cudaGetLastError());
                    const int zm = std::max(1, (m->max_rows_per_image + bt - 1) / bt);
                    dim3 grid2(ni, nf, zm);
                    ProfScope ps(c, PROF_ORB, st);
                    mcl::orb_crosscheck_kernel<<<grid2, bt, 0, st>>>((const uint4*)q->d_desc, q->d_off,
                                                                      (const uint4*)m->d_desc, m->d_src_off, ni, d_mask,
                                                                      c->row_arg.as<int32_t>(), d_counts);
                } else {
                    return fail(MCL_EINVAL, "bad mode %d", mode);
                }
            } else {  // SURF
                if (mode != MCL_KNN2_RATIO) return fail(MCL_EINVAL, "SURF supports mode MCL_KNN2_RATIO only");
                ProfScope ps(c, PROF_SURF, st);
                // few (frame, image, row block) CTAs -- a DOR step -- : cut every map image into row slices, one 128-thread
                // group each, until there are ~8 warps per scheduler; plenty of CTAs: one thread per query row
                const int64_t pairs = (d_mask && c->active_pairs_hint >= 0) ? std::max<int64_t>(1, c->active_pairs_hint) : (int64_t)ni * nf;
                const int64_t blocks128 = pairs * ((max_q + 127) / 128);
                int S = 1;
                while (S < 4 && blocks128 * S * 4 < (int64_t)c->sm_count * 32 && m->max_rows_per_image >= 64 * S) S *= 2;   // 8 slices = 1024 threads = 64 registers: spills
                dim3 grid128(ni, nf, (unsigned)std::max<int64_t>(1, (max_q + 127) / 128));
                const float4* qd = (const float4*)q->d_desc;
                const float4* md = (const float4*)m->d_desc;
                if (S == 4) mcl::surf_knn_sliced_kernel<4><<<grid128, 512, 0, st>>>(qd, q->d_off, md, m->d_src_off, ni, ratio, d_mask, d_counts);
                else if (S == 2) mcl::surf_knn_sliced_kernel<2><<<grid128, 256, 0, st>>>(qd, q->d_off, md, m->d_src_off, ni, ratio, d_mask, d_counts);
                else mcl::surf_knn_kernel<<<grid, bt, 0, st>>>(qd, q->d_off, md, m->d_src_off, ni, ratio, d_mask, d_counts);
            }
            CU(cudaGetLastError());
        }
    }
    if (d_mask) {
        c->launches++;
        mcl::apply_mask_kernel<<<grid_for(n_out, 256, c->sm_count * 4), 256, 0, st>>>(d_counts, d_mask, n_out, fill);
        CU(cudaGetLastError());
    }
    return MCL_OK;
}

}  // namespace

extern "C" int mcl_match_counts_device(mcl_map* m, mcl_queries* q, int mode, double ratio, const uint8_t* d_mask,
                                       int32_t fill_value, int algo, int32_t* d_counts, void* stream) {
    if (!m || !q || !d_counts) return fail(MCL_EINVAL, "mcl_match_counts_device: NULL argument");
    if (m->ctx != q->ctx) return fail(MCL_EINVAL, "map and queries belong to different contexts");
    if (m->kind != q->kind) return fail(MCL_EINVAL, "map kind %d != query kind %d", m->kind, q->kind);
    if (!(ratio > 0.0) || !(ratio <= 1.0)) return fail(MCL_EINVAL, "ratio must be in (0,1]");
    CU(cudaSetDevice(m->ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : m->ctx->stream;
    return match_device(m, q, mode, ratio, d_mask, fill_value, algo, d_counts, st);
}

namespace {
// Host batch of SIFT frames, large enough to be worth pipelining: the frames are cut into 2-3 parts of growing size; part p+1 is
// copied (H2D) and ingested on the copy stream while part p is matched on the main stream.  Returns 1 if the batch is
// too small (caller takes the plain path).
int match_counts_sift_pipelined(mcl_map* m, int n_queries, const int64_t* offs, const void* q_desc, int q_dtype, int mode,
                                double ratio, const uint8_t* mask, int32_t fill_value, int algo, int32_t* out_counts) {
    mcl_ctx* c = m->ctx;
    const int64_t total_rows = offs[n_queries];
    // Two or three parts with growing sizes (1/16, 3/16, 3/4 of the rows): only the first, small part's copy is exposed,
    // and the later, large parts keep the tensor-core launches long (short launches lose a few % to tails and A reloads).
    // All copies are queued at once on a stream that carries only copies (55 GB/s: the whole 210 MB batch of config 3 is
    // on the device after 4 ms, well before the second part's matching ends); each part's memsets and conversion kernel
    // run on the main stream in front of its matching.  (With the conversion kernels on the copy stream every copy
    // waited, in stream order, behind a kernel that could not get an SM while the persistent matching kernel was
    // running: 2.5 ms of gaps per call.)
    int P = total_rows >= 8 * 32768 ? 3 : (total_rows >= 2 * 32768 ? 2 : 1);
    P = std::min(P, n_queries);
    if (P <= 1 || m->n_images == 0) return 1;
    CU(cudaSetDevice(c->device));
    if (!c->copy_stream) CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    while ((int)c->part_events.size() < P) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->part_events.push_back(e);
    }
    std::vector<int> fb(P + 1, n_queries);
    fb[0] = 0;
    {
        const double cut3[2] = {1.0 / 16, 4.0 / 16}, cut2[1] = {1.0 / 8};
        for (int p = 1, f = 0; p < P; ++p) {
            const int64_t target = (int64_t)((P == 3 ? cut3[p - 1] : cut2[p - 1]) * (double)total_rows);
            while (f < n_queries && offs[f] < target) ++f;
            fb[p] = std::max(f, fb[p - 1]);
        }
    }
    const size_t row_b = q_dtype == MCL_F32 ? 512 : 128;
    const int ni = m->n_images;
    const size_t n_out = (size_t)n_queries * ni;
    std::vector<int64_t> region(P + 1, 0);
    for (int p = 0; p < P; ++p) {
        const int64_t rows = offs[fb[p + 1]] - offs[fb[p]];
        region[p + 1] = region[p] + std::max<int64_t>(QROW_PAD, (rows + QROW_PAD - 1) / QROW_PAD * QROW_PAD);
    }
    CU(c->tq_desc.ensure((size_t)region[P] * 128));
    CU(c->tq_norm.ensure((size_t)region[P] * 4));
    CU(c->tq_frame.ensure((size_t)region[P] * 4));
    CU(c->tq_off.ensure((size_t)(n_queries + P) * 8));
    CU(c->stage.ensure((size_t)total_rows * row_b));
    CU(c->tmp_counts.ensure(n_out * 4));
    cudaStream_t st = c->stream, cs = c->copy_stream;
    uint8_t* d_mask = nullptr;
    if (mask) {
        CU(c->tmp_mask.ensure(n_out));
        CU(cudaMemcpyAsync(c->tmp_mask.p, mask, n_out, cudaMemcpyHostToDevice, st));
        d_mask = c->tmp_mask.as<uint8_t>();
    }
    int* bad = c->flags.as<int>() + 2;
    CU(cudaMemsetAsync(bad, 0, 4, st));
    std::vector<mcl_queries> parts(P);
    int rc = MCL_OK;
    // MCL_DEBUG: device timeline of the parts (events on both streams), printed after the call
    const bool dbg = getenv("MCL_DEBUG") != nullptr;
    std::vector<cudaEvent_t> tl;
    auto mark = [&](cudaStream_t s_) {
        if (!dbg) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, s_);
        tl.push_back(e);
    };
    mark(cs);
    for (int p = 0; p < P && rc == MCL_OK; ++p) {
        mcl_queries& q = parts[p];
        const int f0 = fb[p], f1 = fb[p + 1];
        q.ctx = c;
        q.kind = MCL_SIFT;
        q.borrowed = true;
        q.n_queries = f1 - f0;
        q.h_off.resize(q.n_queries + 1);
        for (int f = f0; f <= f1; ++f) q.h_off[f - f0] = offs[f] - offs[f0];
        q.rows = q.h_off.back();
        q.padded_rows = region[p + 1] - region[p];
        q.d_off = c->tq_off.as<int64_t>() + f0 + p;
        q.d_desc = c->tq_desc.as<uint8_t>() + (size_t)region[p] * 128;
        q.d_norm = c->tq_norm.as<int32_t>() + region[p];
        q.d_frame = c->tq_frame.as<int32_t>() + region[p];
        if (q.n_queries == 0) continue;
        // copies on the copy stream (all parts' copies are queued at once and run back to back); the memsets and the
        // conversion kernel of a part on the main stream, right in front of its matching
        CU(cudaMemcpyAsync(q.d_off, q.h_off.data(), (size_t)(q.n_queries + 1) * 8, cudaMemcpyHostToDevice, cs));
        rc = ingest_rows128(c, st, (const uint8_t*)q_desc + (size_t)offs[f0] * row_b, q_dtype, q.rows, q.n_queries, q.d_off,
                            q.d_off, q.padded_rows, (uint8_t*)q.d_desc, q.d_norm, q.d_frame,
                            c->stage.as<uint8_t>() + (size_t)offs[f0] * row_b, false, nullptr, cs, c->part_events[p]);
        if (rc) break;
        mark(cs);
        mark(st);
        rc = match_device(m, &q, mode, ratio, d_mask ? d_mask + (size_t)f0 * ni : nullptr, fill_value, algo,
                          c->tmp_counts.as<int32_t>() + (size_t)f0 * ni, st);
        mark(st);
    }
    if (rc == MCL_OK) {
        cudaError_t e = cudaMemcpyAsync(out_counts, c->tmp_counts.p, n_out * 4, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) rc = fail(MCL_ECUDA, "D2H of counts failed: %s", cudaGetErrorString(e));
    }
    int h_bad = 0;
    cudaStreamSynchronize(cs);
    cudaError_t e1 = cudaStreamSynchronize(st);
    if (dbg && tl.size() == (size_t)(1 + 3 * P)) {
        for (int p = 0; p < P; ++p) {
            float a = 0, b = 0, d = 0;
            cudaEventElapsedTime(&a, tl[0], tl[1 + 3 * p]);       // copy + ingest of the part done (copy stream)
            cudaEventElapsedTime(&b, tl[0], tl[2 + 3 * p]);       // matching of the part may start (main stream)
            cudaEventElapsedTime(&d, tl[0], tl[3 + 3 * p]);       // matching done
            fprintf(stderr, "[mcl] part %d (%d frames): ingested at %.2f ms, matching %.2f -> %.2f ms\n", p, fb[p + 1] - fb[p], a, b, d);
        }
        for (auto e : tl) cudaEventDestroy(e);
    }
    if (rc == MCL_OK && e1 != cudaSuccess) rc = fail(MCL_ECUDA, "stream sync failed: %s", cudaGetErrorString(e1));
    if (rc == MCL_OK && q_dtype == MCL_F32) {
        CU(cudaMemcpy(&h_bad, bad, 4, cudaMemcpyDeviceToHost));
        if (h_bad) rc = fail(MCL_EDATA, "SIFT descriptors must be integer valued in [0,255]");
    }
    return rc;
}
}  // namespace

extern "C" int mcl_match_counts(mcl_map* m, int n_queries, const int64_t* q_row_offsets, const void* q_desc, int q_dtype,
                                int mode, double ratio, const uint8_t* mask, int32_t fill_value, int algo,
                                int32_t* out_counts) {
    if (!m || !out_counts) return fail(MCL_EINVAL, "mcl_match_counts: NULL argument");
    if (!(ratio > 0.0) || !(ratio <= 1.0)) return fail(MCL_EINVAL, "ratio must be in (0,1]");
    mcl_ctx* c = m->ctx;
    if (m->kind == MCL_SIFT && n_queries > 1 && q_row_offsets && q_desc && (q_dtype == MCL_U8 || q_dtype == MCL_F32) &&
        check_offsets(q_row_offsets, n_queries, "mcl_match_counts") == MCL_OK && n_queries <= 65535) {
        const int r = match_counts_sift_pipelined(m, n_queries, q_row_offsets, q_desc, q_dtype, mode, ratio, mask, fill_value,
                                                  algo, out_counts);
        if (r != 1) return r;
    }
    mcl_queries* q = nullptr;
    int rc = queries_build(c, m->kind, n_queries, q_row_offsets, q_desc, q_dtype, true, &q);
    if (rc) return rc;
    const size_t n_out = (size_t)n_queries * m->n_images;
    cudaStream_t st = c->stream;
    auto body = [&]() -> int {
        if (n_out == 0) return MCL_OK;
        CU(c->tmp_counts.ensure(n_out * 4));
        uint8_t* d_mask = nullptr;
        if (mask) {
            CU(c->tmp_mask.ensure(n_out));
            CU(cudaMemcpyAsync(c->tmp_mask.p, mask, n_out, cudaMemcpyHostToDevice, st));
            d_mask = c->tmp_mask.as<uint8_t>();
        }
        c->active_pairs_hint = -1;
        if (mask) {
            int64_t act = 0;
            for (size_t i = 0; i < n_out; ++i) act += mask[i] != 0;
            c->active_pairs_hint = act;
        }
        int r = match_device(m, q, mode, ratio, d_mask, fill_value, algo, c->tmp_counts.as<int32_t>(), st);
        c->active_pairs_hint = -1;
        if (r) return r;
        CU(cudaMemcpyAsync(out_counts, c->tmp_counts.p, n_out * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        return MCL_OK;
    };
    rc = body();
    mcl_queries_destroy(q);
    return rc;
}

// ================================================================================================
// bag of words
// ================================================================================================
namespace {
void bow_free_tc(mcl_bow* b) {
    cudaFree(b->d_tile_off);
    cudaFree(b->d_tiles);
    cudaFree(b->d_rec);
    cudaFree(b->d_key_c);
    for (auto& sp : b->splits) cudaFree(sp.d);
    b->d_tile_off = nullptr; b->d_tiles = nullptr; b->d_rec = nullptr; b->d_key_c = nullptr;
    b->splits.clear();
    b->cap_tiles = b->cap_loc = 0;
    b->tc_ok = false;
}

// (re)builds the tensor-core structures of an index from its float32 code book (d_voc, d_voc_off, h_voc_off, n_loc)
int bow_build_tc(mcl_bow* b, cudaStream_t st) {
    using namespace mcl::tc9;
    mcl_ctx* c = b->ctx;
    const int L = b->n_loc;
    b->h_tile_off.resize(L + 1);
    int32_t acc = 0;
    for (int l = 0; l < L; ++l) {
        b->h_tile_off[l] = acc;
        acc += (int32_t)((b->h_voc_off[l + 1] - b->h_voc_off[l] + TILE_N - 1) / TILE_N);
    }
    b->h_tile_off[L] = acc;
    const bool reuse = b->d_tiles && acc <= b->cap_tiles && L == b->cap_loc;   // a rebuild that fits (k-means: k only shrinks)
    if (!reuse) {
        bow_free_tc(b);
        CU(cudaMalloc(&b->d_tile_off, (size_t)(L + 1) * 4));
        CU(cudaMalloc(&b->d_tiles, std::max<size_t>((size_t)acc * Shape<KIND_BOW>::TILE_BYTES, 256)));
        CU(cudaMalloc(&b->d_rec, std::max<size_t>((size_t)acc * sizeof(TileRec), 256)));
        CU(cudaMalloc(&b->d_key_c, (size_t)L * 4));
        for (int ns : {1, 2, 4, 8}) {
            mcl_bow::Split sp;
            sp.n_splits = ns;
            CU(cudaMalloc(&sp.d, std::max<size_t>((size_t)ns * L * sizeof(mcl::sifttc3::Chunk), 16)));
            b->splits.push_back(sp);
        }
        b->cap_tiles = acc;
        b->cap_loc = L;
    }
    CU(cudaMemcpyAsync(b->d_tile_off, b->h_tile_off.data(), (size_t)(L + 1) * 4, cudaMemcpyHostToDevice, st));
    int* bad = c->flags.as<int>() + 6;
    CU(cudaMemsetAsync(bad, 0, 4, st));
    {
        ProfScope ps(c, PROF_PREP, st);
        bow_key_c_kernel<<<L, 256, 0, st>>>(b->d_voc, b->d_voc_off, b->d_key_c);
        bow_build_tiles_kernel<<<acc, 256, 0, st>>>(b->d_voc, b->d_voc_off, L, b->d_tile_off, b->d_key_c, b->d_tiles, b->d_rec, bad);
    }
    c->launches++;
    CU(cudaGetLastError());
    std::vector<std::vector<mcl::sifttc3::Chunk>> tables;   // kept alive until the copies below have been consumed
    for (auto& s : b->splits) {
        const int ns = s.n_splits;
        tables.emplace_back();
        auto& ch = tables.back();
        for (int sp = 0; sp < ns; ++sp)
            for (int l = 0; l < L; ++l) {
                const int t0 = b->h_tile_off[l], nt = b->h_tile_off[l + 1] - t0, per = (nt + ns - 1) / ns;
                mcl::sifttc3::Chunk k;
                k.tile_begin = t0 + std::min(nt, sp * per);
                k.tile_end = t0 + std::min(nt, (sp + 1) * per);
                k.img_begin = l;    // location
                k.img_end = sp;     // range index: where the range's records go
                if (k.tile_end > k.tile_begin) ch.push_back(k);
            }
        s.n_chunks = (int)ch.size();
        if (!ch.empty()) CU(cudaMemcpyAsync(s.d, ch.data(), ch.size() * sizeof(mcl::sifttc3::Chunk), cudaMemcpyHostToDevice, st));
    }
    int h_bad = 0;
    CU(cudaMemcpyAsync(&h_bad, bad, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    b->tc_ok = h_bad == 0;   // centroids outside [0, 256): the float32 SIMT kernel is used instead
    return MCL_OK;
}
}  // namespace

extern "C" void mcl_bow_destroy(mcl_bow* b) {
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    cudaFree(b->d_voc);
    cudaFree(b->d_idf);
    cudaFree(b->d_imf);
    cudaFree(b->d_voc_off);
    cudaFree(b->d_img_off);
    bow_free_tc(b);
    delete b;
}

extern "C" int mcl_bow_create(mcl_ctx* c, int n_locations, const int64_t* voc_row_offsets, const float* voc,
                              const float* idf, const int64_t* img_offsets, const float* im_features, int W,
                              mcl_bow** out) {
    if (!c || !out || !voc || !idf || !im_features) return fail(MCL_EINVAL, "mcl_bow_create: NULL argument");
    *out = nullptr;
    if (n_locations <= 0 || W <= 0) return fail(MCL_EINVAL, "mcl_bow_create: n_locations and W must be positive");
    if ((size_t)W * 12 > 200 * 1024) return fail(MCL_EINVAL, "mcl_bow_create: W=%d does not fit the shared-memory histogram + word list", W);
    int rc = check_offsets(voc_row_offsets, n_locations, "mcl_bow_create(voc)");
    if (rc) return rc;
    rc = check_offsets(img_offsets, n_locations, "mcl_bow_create(images)");
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    mcl_bow* b = new (std::nothrow) mcl_bow();
    if (!b) return fail(MCL_EINTERNAL, "out of host memory");
    b->ctx = c;
    b->n_loc = n_locations;
    b->W = W;
    b->h_voc_off.assign(voc_row_offsets, voc_row_offsets + n_locations + 1);
    b->h_img_off.assign(img_offsets, img_offsets + n_locations + 1);
    b->voc_rows = voc_row_offsets[n_locations];
    b->total_images = img_offsets[n_locations];
    for (int l = 0; l < n_locations; ++l) {
        const int64_t nw = voc_row_offsets[l + 1] - voc_row_offsets[l];
        if (nw > W) { delete b; return fail(MCL_EINVAL, "location %d has %lld words > W=%d", l, (long long)nw, W); }
        if (nw <= 0) { delete b; return fail(MCL_EINVAL, "location %d has an empty vocabulary", l); }
        b->max_words = (int)std::max<int64_t>(b->max_words, nw);
    }
    cudaStream_t st = c->stream;
#define BCU(expr)                                                                                                 \
    do {                                                                                                          \
        cudaError_t e_ = (expr);                                                                                  \
        if (e_ != cudaSuccess) {                                                                                  \
            mcl_bow_destroy(b);                                                                                   \
            return fail(MCL_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__);   \
        }                                                                                                         \
    } while (0)
    BCU(cudaMalloc(&b->d_voc, (size_t)b->voc_rows * 512));
    BCU(cudaMalloc(&b->d_idf, (size_t)n_locations * W * 4));
    BCU(cudaMalloc(&b->d_imf, std::max<size_t>((size_t)b->total_images * W * 4, 256)));
    BCU(cudaMalloc(&b->d_voc_off, (size_t)(n_locations + 1) * 8));
    BCU(cudaMalloc(&b->d_img_off, (size_t)(n_locations + 1) * 8));
    BCU(cudaMemcpyAsync(b->d_voc, voc, (size_t)b->voc_rows * 512, cudaMemcpyHostToDevice, st));
    BCU(cudaMemcpyAsync(b->d_idf, idf, (size_t)n_locations * W * 4, cudaMemcpyHostToDevice, st));
    if (b->total_images) {
        // the tf-idf rows are kept TRANSPOSED per location, (W, N_l): the score kernels read one contiguous row per query word
        float* d_tmp = nullptr;
        BCU(cudaMalloc(&d_tmp, (size_t)b->total_images * W * 4));
        cudaError_t e1 = cudaMemcpyAsync(d_tmp, im_features, (size_t)b->total_images * W * 4, cudaMemcpyHostToDevice, st);
        cudaError_t e2 = cudaMemcpyAsync(b->d_img_off, b->h_img_off.data(), (size_t)(n_locations + 1) * 8, cudaMemcpyHostToDevice, st);
        c->launches++;
        mcl::bow_transpose_imf_kernel<<<dim3((W + 31) / 32, n_locations), dim3(32, 8), 0, st>>>(d_tmp, b->d_img_off, W, b->d_imf);
        cudaError_t e3 = cudaStreamSynchronize(st);
        cudaFree(d_tmp);
        BCU(e1);
        BCU(e2);
        BCU(e3);
    }
    BCU(cudaMemcpyAsync(b->d_voc_off, b->h_voc_off.data(), (size_t)(n_locations + 1) * 8, cudaMemcpyHostToDevice, st));
    BCU(cudaMemcpyAsync(b->d_img_off, b->h_img_off.data(), (size_t)(n_locations + 1) * 8, cudaMemcpyHostToDevice, st));
    if (W * 12 > c->bow_score_smem) {   // only ever raised: a smaller index created later must not lower it under a live larger one
        BCU(cudaFuncSetAttribute(mcl::bow_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, W * 12));
        c->bow_score_smem = W * 12;
    }
    BCU(cudaStreamSynchronize(st));
    {   // fp16 code-book tiles for the tensor-core word assignment
        const int r = bow_build_tc(b, st);
        if (r) { mcl_bow_destroy(b); return r; }
    }
    BCU(cudaStreamSynchronize(st));
#undef BCU
    *out = b;
    return MCL_OK;
}

namespace {
// Visual-word assignment of n_q query rows (float32, (n_q,128) on the device) against every location of `b`:
// fills ctx->bow_best[loc * n_q + row] with (float32 distance bits << 32 | word) or (0 << 32 | word).
int bow_assign(mcl_bow* b, const float* d_q, int n_q, cudaStream_t st) {
    mcl_ctx* c = b->ctx;
    const size_t nbest = (size_t)b->n_loc * std::max(1, n_q);
    CU(c->bow_best.ensure(nbest * 8));
    CU(cudaMemsetAsync(c->bow_best.p, 0xFF, nbest * 8, st));
    if (n_q > 0) {
        const int* only_if = nullptr;
        if (b->tc_ok && c->opt[2] != 1) {
            // ---- tensor cores: fp16 copy of the query rows, approximate keys per 32-word group, float32 resolve
            using namespace mcl::tc9;
            using namespace mcl::bowtc;
            {
                const int rc9 = tc9_prepare(c);
                if (rc9) return rc9;
            }
            const int n_qblocks = (n_q + QBLK - 1) / QBLK;
            const int64_t padded = (int64_t)n_qblocks * QBLK;
            const int subs_per_loc = (b->max_words + GROUP - 1) / GROUP * SUBS;   // 8-word sub-blocks per location
            const int64_t n_buckets = (int64_t)b->n_loc * subs_per_loc;
            // work items = (code-book range, location, pair of query blocks): the number of ranges per location with the
            // smallest estimated makespan (rounds of items over the co-resident CTA pairs x (tiles per range + pipeline refill))
            const int n_qgroups = (n_qblocks + 1) / 2;
            const mcl_bow::Split* sp = &b->splits.front();
            {
                const double tiles_per_loc = (double)b->h_tile_off[b->n_loc] / b->n_loc;
                const int slots = std::max(1, c->tc9_clusters[1] > 0 ? c->tc9_clusters[1] : c->sm_count / 2);
                double best_cost = 1e300;
                for (const auto& cand : b->splits) {
                    const int64_t items = (int64_t)cand.n_chunks * n_qgroups;
                    const double cost = (double)((items + slots - 1) / slots) * (tiles_per_loc / cand.n_splits + 1.5);
                    if (cost < best_cost) { best_cost = cost; sp = &cand; }
                }
            }
            const int n_splits = sp->n_splits;
            CU(c->bow_q8.ensure((size_t)padded * Shape<KIND_BOW>::ROW_BYTES));
            CU(c->bow_qaux.ensure((size_t)padded * 4));
            if (nbest >= ((size_t)1 << 29)) return fail(MCL_EINVAL, "bow: too many (location, row) pairs in one call");
            CU(c->bow_tc_out.ensure((size_t)n_splits * nbest * 3 * sizeof(int4)));
            // sort scratch: sorted[12 nbest] (int2: a row has at most 3 groups x 4 sub-blocks) | job_sel[nbest][3] | queue[nbest] |
            // bucket_cnt[n_buckets]
            CU(c->bow_sort.ensure((28 * nbest + (size_t)n_buckets) * 4));
            int2* sorted = c->bow_sort.as<int2>();
            int* job_grp = reinterpret_cast<int*>(sorted + 12 * nbest);
            int* queue = job_grp + 3 * nbest;
            int* bucket = queue + nbest;
            int* bad = c->flags.as<int>() + 6;
            unsigned int* state = c->flags.as<unsigned int>() + 8;   // [8] queued rows, [9] SIMT flag, [10] rescans
            CU(cudaMemsetAsync(c->bow_q8.p, 0, (size_t)padded * Shape<KIND_BOW>::ROW_BYTES, st));
            CU(cudaMemsetAsync(c->bow_qaux.p, 0xFF, (size_t)padded * 4, st));
            CU(cudaMemsetAsync(c->bow_tc_out.p, 0, (size_t)n_splits * nbest * 3 * sizeof(int4), st));
            CU(cudaMemsetAsync(bucket, 0, (size_t)n_buckets * 4, st));
            CU(cudaMemsetAsync(bad, 0, 4, st));
            CU(cudaMemsetAsync(state, 0, 8, st));
            CU(cudaMemsetAsync(state + 3, 0, 4, st));
            {
                ProfScope ps(c, PROF_PREP, st);
                bow_query_prep_kernel<<<grid_for((int64_t)n_q * 32, 256, c->sm_count * 8), 256, 0, st>>>(
                    d_q, n_q, c->bow_q8.as<uint8_t>(), c->bow_qaux.as<int32_t>(), bad);
            }
            CU(cudaGetLastError());
            Params p = {};
            p.q_rows = c->bow_q8.as<uint8_t>();
            p.q_aux = c->bow_qaux.as<int32_t>();
            p.m_tiles = b->d_tiles;
            p.m_rec = b->d_rec;
            p.chunks = sp->d;
            p.n_qblocks = n_qblocks;
            p.n_chunks = sp->n_chunks;
            p.out = c->bow_tc_out.as<int4>();
            p.n_q = n_q;
            p.n_loc = b->n_loc;
            p.meta_cap = META_CAP;
            // CTA pairs share the code-book tiles (multicast) when there are query blocks to pair; option 5 = 1 turns it off
            int cl = 1;
            if (c->opt[5] == 2 || (c->opt[5] == 0 && n_qblocks >= 2 && (n_qblocks % 2 == 0 || n_qblocks >= 16))) cl = 2;
            if (cl == 2 && c->tc9_clusters[1] < 1) cl = 1;
            const int64_t n_items = (int64_t)((n_qblocks + cl - 1) / cl) * sp->n_chunks;
            const int grid9 = cl == 2 ? 2 * (int)std::min<int64_t>(n_items, c->tc9_clusters[1]) : (int)std::min<int64_t>(n_items, c->sm_count);
            #ifdef MCL_TC_DIAG
        static const bool diag9 = getenv("MCL_TC9_DIAG") != nullptr;
#else
        constexpr bool diag9 = false;
#endif
            if (diag9) {
                CU(c->diag9.ensure((size_t)grid9 * 5 * sizeof(long long)));
                CU(cudaMemsetAsync(c->diag9.p, 0, (size_t)grid9 * 5 * sizeof(long long), st));
                p.diag = c->diag9.as<long long>();
            }
            {
                ProfScope ps(c, PROF_VQ_TC, st);
                constexpr size_t smem = sizeof(Smem<KIND_BOW>);
                if (cl == 2) CU(launch_cluster(tc9_kernel<KIND_BOW, 2>, grid9, THREADS, smem, st, 2, p));
                else tc9_kernel<KIND_BOW, 1><<<grid9, THREADS, smem, st>>>(p);
            }
            if (diag9) {
                const int r = tc9_diag_report(c, st, "BOW", p.diag, grid9, sp->n_chunks, (int)((n_qblocks + cl - 1) / cl));
                if (r) return r;
            }
            CU(cudaGetLastError());
            {
                ProfScope ps(c, PROF_VQ_RESOLVE, st);
                const int g = grid_for((int64_t)nbest, 256, c->sm_count * 8);
                bow_merge_kernel<<<g, 256, 0, st>>>(c->bow_tc_out.as<int4>(), n_splits, n_q, b->n_loc, subs_per_loc, bad, job_grp,
                                                    bucket, queue, state);
                bow_scan_kernel<<<1, 1024, 0, st>>>(bucket, (int)n_buckets, state);
                bow_scatter_kernel<<<g, 256, 0, st>>>(job_grp, n_q, (int64_t)nbest, subs_per_loc, bucket, sorted);
                // entries per warp: enough warps for ~16 per SM, runs of 8..64 entries
                const int run_len = (int)std::max<int64_t>(8, std::min<int64_t>(64, (int64_t)nbest / ((int64_t)c->sm_count * 16)));
                bow_resolve_kernel<<<grid_for((2 * (int64_t)nbest + run_len - 1) / run_len * 32, 128, c->sm_count * 32), 128, 0, st>>>(
                    sorted, state, run_len, d_q, n_q, subs_per_loc, b->d_voc, b->d_voc_off, c->bow_best.as<unsigned long long>());
            }
            {
                ProfScope ps(c, PROF_VQ_RESCAN, st);
                bow_decide_kernel<<<1, 1, 0, st>>>(bad, (int64_t)nbest, state);
                bow_rescan_kernel<<<c->sm_count * 8, 256, 0, st>>>(queue, d_q, n_q, b->d_voc, b->d_voc_off,
                                                                  c->bow_best.as<unsigned long long>(), state);
            }
            c->launches += 5;
            CU(cudaGetLastError());
            only_if = reinterpret_cast<const int*>(state + 1);   // many ambiguous rows or non-integer queries: float32 SIMT pass
        }
        const int gx = (b->max_words + mcl::VQ_WORDS - 1) / mcl::VQ_WORDS, gy = (n_q + 127) / 128, gz = b->n_loc;
        const int64_t n_work = (int64_t)gx * gy * gz;
        // on its own the kernel gets one block per work item; as the (usually idle) fallback of the tensor-core path a
        // small grid walks the work space
        const int64_t blocks = only_if ? std::min<int64_t>(n_work, (int64_t)c->sm_count * 16) : std::min<int64_t>(n_work, (int64_t)1 << 30);
        ProfScope ps(c, PROF_VQ, st);
        mcl::bow_vq_kernel<<<(unsigned)blocks, 128, 0, st>>>((const float4*)d_q, n_q, (const float4*)b->d_voc, b->d_voc_off, only_if,
                                                             gx, gy, gz, c->bow_best.as<unsigned long long>());
    }
    CU(cudaGetLastError());
    return MCL_OK;
}
}  // namespace

extern "C" int mcl_bow_score_device(mcl_bow* b, mcl_queries* q, int32_t* d_words, float* d_scores, void* stream) {
    if (!b || !q || !d_scores) return fail(MCL_EINVAL, "mcl_bow_score_device: NULL argument");
    if (q->kind != MCL_BOWQ) return fail(MCL_EINVAL, "mcl_bow_score_device: queries must be of kind MCL_BOWQ");
    if (b->ctx != q->ctx) return fail(MCL_EINVAL, "bow index and queries belong to different contexts");
    mcl_ctx* c = b->ctx;
    CU(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const int n_q = (int)q->rows;
    const int nf = q->n_queries;
    if (nf == 0) return MCL_OK;
    if (nf > 65535) return fail(MCL_EINVAL, "at most 65535 query frames per call");
    {
        const int r = bow_assign(b, (const float*)q->d_desc, n_q, st);
        if (r) return r;
    }
    {
        dim3 grid(b->n_loc, nf);
        int64_t max_rows = 1;
        for (int f = 0; f < nf; ++f) max_rows = std::max(max_rows, q->h_off[f + 1] - q->h_off[f]);
        const int list_cap = (int)std::min<int64_t>(b->W, max_rows);   // distinct words of a frame <= its rows
        ProfScope ps(c, PROF_SCORE, st);
        if (max_rows <= mcl::BOW_SPARSE_ROWS)
            mcl::bow_score_sparse_kernel<<<grid, 256, 0, st>>>(c->bow_best.as<unsigned long long>(), n_q, q->d_off, b->d_idf, b->d_imf,
                                                               b->d_img_off, b->W, (int)b->total_images, d_words, d_scores);
        else
            mcl::bow_score_kernel<<<grid, 256, (size_t)b->W * 4 + (size_t)list_cap * 8, st>>>(
                c->bow_best.as<unsigned long long>(), n_q, q->d_off, b->d_idf, b->d_imf, b->d_img_off, b->W,
                (int)b->total_images, list_cap, d_words, d_scores);
    }
    CU(cudaGetLastError());
    return MCL_OK;
}

extern "C" int mcl_bow_score(mcl_bow* b, int n_queries, const int64_t* q_row_offsets, const void* q_desc, int q_dtype,
                             int32_t* out_words, float* out_scores) {
    if (!b || !out_scores) return fail(MCL_EINVAL, "mcl_bow_score: NULL argument");
    mcl_ctx* c = b->ctx;
    mcl_queries* q = nullptr;
    int rc = queries_build(c, MCL_BOWQ, n_queries, q_row_offsets, q_desc, q_dtype, true, &q);
    if (rc) return rc;
    cudaStream_t st = c->stream;
    auto body = [&]() -> int {
        const size_t n_scores = (size_t)n_queries * b->total_images;
        const size_t n_words = (size_t)b->n_loc * q->rows;
        if (n_queries == 0) return MCL_OK;
        CU(c->tmp_scores.ensure(std::max<size_t>(n_scores * 4, 256)));
        int32_t* d_words = nullptr;
        if (out_words && n_words) {
            CU(c->tmp_words.ensure(n_words * 4));
            d_words = c->tmp_words.as<int32_t>();
        }
        int r = mcl_bow_score_device(b, q, d_words, c->tmp_scores.as<float>(), st);
        if (r) return r;
        if (n_scores) CU(cudaMemcpyAsync(out_scores, c->tmp_scores.p, n_scores * 4, cudaMemcpyDeviceToHost, st));
        if (d_words) CU(cudaMemcpyAsync(out_words, d_words, n_words * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        return MCL_OK;
    };
    rc = body();
    mcl_queries_destroy(q);
    return rc;
}

// ================================================================================================
// k-means code book training (BOW index build)
// ================================================================================================
extern "C" int mcl_kmeans(mcl_ctx* c, int64_t n, const void* obs, int obs_dtype, int k, const float* guess, double thresh,
                          int max_iter, float* out_voc, int* out_k, double* out_distortion, int* out_iters) {
    if (!c || !obs || !guess || !out_voc || !out_k) return fail(MCL_EINVAL, "mcl_kmeans: NULL argument");
    if (n <= 0 || k <= 0 || n > ((int64_t)1 << 30)) return fail(MCL_EINVAL, "mcl_kmeans: bad n or k");
    if (obs_dtype != MCL_U8 && obs_dtype != MCL_F32) return fail(MCL_EINVAL, "mcl_kmeans: bad dtype");
    if (max_iter <= 0) max_iter = INT_MAX;   // scipy's loop has no iteration cap
    CU(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    float* d_obs = nullptr;
    float* d_voc[2] = {nullptr, nullptr};
    float* d_sums = nullptr;
    int* d_counts = nullptr;
    int64_t* d_voc_off = nullptr;
    double* d_sum = nullptr;   // [0] distortion sum; the following int is k_new
    mcl_bow tmp;   // a one-location "index" over the current code book, for bow_assign
    auto cleanup = [&]() {   // idempotent: an error inside assign_and_distort frees here and the caller frees again
        cudaStreamSynchronize(st);
        void** ptrs[] = {(void**)&d_obs, (void**)&d_voc[0], (void**)&d_voc[1], (void**)&d_sums, (void**)&d_counts,
                         (void**)&d_voc_off, (void**)&d_sum};
        for (void** p : ptrs) { if (*p) cudaFree(*p); *p = nullptr; }
        bow_free_tc(&tmp);
    };
#define KCU(expr)                                                                                                \
    do {                                                                                                          \
        cudaError_t e_ = (expr);                                                                                  \
        if (e_ != cudaSuccess) {                                                                                  \
            cleanup();                                                                                            \
            return fail(MCL_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__);   \
        }                                                                                                         \
    } while (0)
    KCU(cudaMalloc(&d_obs, (size_t)n * 512));
    KCU(cudaMalloc(&d_voc[0], (size_t)k * 512));
    KCU(cudaMalloc(&d_voc[1], (size_t)k * 512));
    KCU(cudaMalloc(&d_sums, (size_t)k * 512));
    KCU(cudaMalloc(&d_counts, (size_t)k * 4));
    KCU(cudaMalloc(&d_voc_off, 16));
    KCU(cudaMalloc(&d_sum, 16));
    if (obs_dtype == MCL_F32) {
        KCU(cudaMemcpyAsync(d_obs, obs, (size_t)n * 512, cudaMemcpyHostToDevice, st));
    } else {
        KCU(c->stage.ensure((size_t)n * 128));
        KCU(cudaMemcpyAsync(c->stage.p, obs, (size_t)n * 128, cudaMemcpyHostToDevice, st));
        c->launches++;
        mcl::u8_to_f32_kernel<<<grid_for(n * 128, 256, c->sm_count * 8), 256, 0, st>>>(c->stage.as<uint8_t>(), d_obs, n * 128);
    }
    KCU(cudaMemcpyAsync(d_voc[0], guess, (size_t)k * 512, cudaMemcpyHostToDevice, st));
    tmp.ctx = c;
    tmp.n_loc = 1;
    tmp.W = k;
    tmp.d_voc_off = d_voc_off;
    int cur = 0, k_cur = k, iters = 0;
    double prev = INFINITY, avg = 0.0;
    int rc = MCL_OK;
    auto assign_and_distort = [&]() -> int {
        const int64_t h_off[2] = {0, k_cur};
        KCU(cudaMemcpyAsync(d_voc_off, h_off, 16, cudaMemcpyHostToDevice, st));
        KCU(cudaStreamSynchronize(st));
        tmp.d_voc = d_voc[cur];
        tmp.max_words = k_cur;
        tmp.voc_rows = k_cur;
        tmp.h_voc_off.assign(h_off, h_off + 2);
        {
            const int rb = bow_build_tc(&tmp, st);   // fp16 tiles of the current code book
            if (rb) return rb;
        }
        const int r = bow_assign(&tmp, d_obs, (int)n, st);
        if (r) return r;
        KCU(cudaMemsetAsync(d_sum, 0, 16, st));
        c->launches++;
        mcl::kmeans_distort_kernel<<<c->sm_count * 8, 256, 0, st>>>(d_obs, n, d_voc[cur], c->bow_best.as<unsigned long long>(), d_sum);
        double h_sum = 0.0;
        KCU(cudaMemcpyAsync(&h_sum, d_sum, 8, cudaMemcpyDeviceToHost, st));
        KCU(cudaStreamSynchronize(st));
        avg = h_sum / (double)n;
        return MCL_OK;
    };
    // scipy _kmeans: while diff > thresh: assign, avg distortion, update, diff = |previous avg - avg|
    double diff = INFINITY;
    while (diff > thresh && iters < max_iter) {
        rc = assign_and_distort();
        if (rc) break;
        KCU(cudaMemsetAsync(d_sums, 0, (size_t)k_cur * 512, st));
        KCU(cudaMemsetAsync(d_counts, 0, (size_t)k_cur * 4, st));
        c->launches += 2;
        mcl::kmeans_accum_kernel<<<c->sm_count * 8, 256, 0, st>>>(d_obs, n, c->bow_best.as<unsigned long long>(), d_sums, d_counts);
        int* d_knew = reinterpret_cast<int*>(d_sum + 1);
        mcl::kmeans_update_kernel<<<1, 1024, 0, st>>>(k_cur, d_sums, d_counts, d_voc[cur ^ 1], d_knew);
        int h_knew = 0;
        KCU(cudaMemcpyAsync(&h_knew, d_knew, 4, cudaMemcpyDeviceToHost, st));
        KCU(cudaStreamSynchronize(st));
        diff = std::fabs(prev - avg);
        prev = avg;
        cur ^= 1;
        k_cur = h_knew;
        ++iters;
        if (k_cur <= 0) { rc = fail(MCL_EDATA, "mcl_kmeans: every cluster became empty"); break; }
    }
    if (rc == MCL_OK && iters == 0) rc = assign_and_distort();   // thresh = inf: scipy would not iterate either; report the guess's distortion
    if (rc == MCL_OK) {
        KCU(cudaMemcpyAsync(out_voc, d_voc[cur], (size_t)k_cur * 512, cudaMemcpyDeviceToHost, st));
        KCU(cudaStreamSynchronize(st));
        *out_k = k_cur;
        // scipy returns prev_avg_dists[1]: the average distortion of the LAST assignment, i.e. against the code book before
        // its final update (scipy/cluster/vq/_vq_impl.py _kmeans)
        if (out_distortion) *out_distortion = avg;
        if (out_iters) *out_iters = iters;
    }
    tmp.d_voc = nullptr; tmp.d_voc_off = nullptr;
    cleanup();
#undef KCU
    return rc;
}

// ================================================================================================
// small stages
// ================================================================================================
namespace {
// the recursion of numpy's pairwise sum over n elements starting at lo; returns a reference to the node holding the
// range's sum: a leaf index (>= 0) or -1 - (index of an internal node in `nodes`)
struct LapNode { int l, r, height; };
int lap_build(int lo, int n, std::vector<int>& leaf_lo, std::vector<int>& leaf_n, std::vector<LapNode>& nodes, int& height) {
    if (n <= 128) {
        leaf_lo.push_back(lo);
        leaf_n.push_back(n);
        height = 0;
        return (int)leaf_lo.size() - 1;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    int hl = 0, hr = 0;
    const int l = lap_build(lo, n2, leaf_lo, leaf_n, nodes, hl);
    const int r = lap_build(lo + n2, n - n2, leaf_lo, leaf_n, nodes, hr);
    height = std::max(hl, hr) + 1;
    nodes.push_back({l, r, height});
    return -1 - ((int)nodes.size() - 1);
}

int lap_plan_get(mcl_ctx* c, int64_t N, const LapPlan** out) {
    auto it = c->lap_plans.find(N);
    if (it != c->lap_plans.end()) { *out = &it->second; return MCL_OK; }
    if (c->lap_plans.size() >= 32) {   // sizes come and go: drop the tables rather than grow without bound
        CU(cudaStreamSynchronize(c->stream));
        for (auto& kv : c->lap_plans) cudaFree(kv.second.d_tab);
        c->lap_plans.clear();
    }
    std::vector<int> leaf_lo, leaf_n;
    std::vector<LapNode> nodes;
    int height = 0;
    lap_build(0, (int)N, leaf_lo, leaf_n, nodes, height);
    // internal nodes ordered by height (children before parents; the root last), references remapped to value indices
    std::vector<int> order(nodes.size());
    for (size_t i = 0; i < nodes.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return nodes[a].height < nodes[b].height; });
    std::vector<int> pos(nodes.size());
    for (size_t i = 0; i < order.size(); ++i) pos[order[i]] = (int)i;
    LapPlan pl;
    pl.n_leaves = (int)leaf_lo.size();
    pl.n_nodes = (int)nodes.size();
    pl.n_levels = height;
    std::vector<int> tab;
    tab.insert(tab.end(), leaf_lo.begin(), leaf_lo.end());
    tab.insert(tab.end(), leaf_n.begin(), leaf_n.end());
    auto ref = [&](int r) { return r >= 0 ? r : pl.n_leaves + pos[-1 - r]; };
    for (int i : order) tab.push_back(ref(nodes[i].l));
    for (int i : order) tab.push_back(ref(nodes[i].r));
    std::vector<int> level_off(height + 1, 0);   // nodes of height h sit at [level_off[h - 1], level_off[h])
    for (int i : order) level_off[nodes[i].height]++;
    for (int h = 1; h <= height; ++h) level_off[h] += level_off[h - 1];
    tab.insert(tab.end(), level_off.begin(), level_off.end());
    CU(cudaMalloc((void**)&pl.d_tab, std::max<size_t>(tab.size() * 4, 16)));
    CU(cudaMemcpy(pl.d_tab, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
    *out = &c->lap_plans.emplace(N, pl).first->second;
    return MCL_OK;
}

int laplacian_common(mcl_ctx* c, int n, const uint8_t* const* imgs, const int* h, const int* w, const int* stride,
                     int64_t* sums_out, double* var_out, const uint8_t* packed = nullptr) {
    if (!c || (n > 0 && ((!imgs && !packed) || !h || !w))) return fail(MCL_EINVAL, "mcl_laplacian: NULL argument");
    if (n <= 0) return MCL_OK;
    if (n > 65535) return fail(MCL_EINVAL, "at most 65535 images per call");
    CU(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    std::vector<int64_t> off(n + 1, 0);
    std::vector<int> pitch(n);
    int max_h = 0, max_w = 0;
    for (int i = 0; i < n; ++i) {
        if (h[i] <= 0 || w[i] <= 0 || (!packed && !imgs[i])) return fail(MCL_EINVAL, "image %d: bad size or NULL pointer", i);
        if (stride && stride[i] < w[i]) return fail(MCL_EINVAL, "image %d: stride < width", i);
        pitch[i] = (w[i] + 15) & ~15;   // rows start on 16-byte boundaries: one aligned 128-bit load per 16 pixels
        off[i + 1] = off[i] + (int64_t)h[i] * pitch[i];
        max_h = std::max(max_h, h[i]);
        max_w = std::max(max_w, w[i]);
    }
    CU(c->lap_pix.ensure((size_t)off[n] + 32));
    // meta: off (n+1 int64) | h (n int) | w (n int) | pitch (n int) | sums (2n int64) | var (n double)
    const size_t o_h = (size_t)(n + 1) * 8, o_w = o_h + (size_t)n * 4, o_p = o_w + (size_t)n * 4,
                 o_s = (o_p + (size_t)n * 4 + 7) & ~(size_t)7, o_v = o_s + (size_t)n * 16, total = o_v + (size_t)n * 8;
    CU(c->lap_meta.ensure(total));
    uint8_t* mb = c->lap_meta.as<uint8_t>();
    if (packed) {   // the caller laid the images out exactly like the device buffer (16-byte row pitch, back to back): ONE copy
        CU(cudaMemcpyAsync(c->lap_pix.p, packed, (size_t)off[n], cudaMemcpyHostToDevice, st));
    } else {
        for (int i = 0; i < n; ++i) {
            const int s = stride ? stride[i] : w[i];
            CU(cudaMemcpy2DAsync(c->lap_pix.as<uint8_t>() + off[i], (size_t)pitch[i], imgs[i], (size_t)s, (size_t)w[i], (size_t)h[i],
                                 cudaMemcpyHostToDevice, st));
        }
    }
    CU(cudaMemcpyAsync(mb, off.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(mb + o_h, h, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(mb + o_w, w, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(mb + o_p, pitch.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(mb + o_s, 0, (size_t)n * 16, st));
    if (sums_out) {   // sum Lb^2 (exact integers); the variance below needs only sum L, which the finish kernel gets from the border
        const int64_t threads = (int64_t)((max_w + mcl::LAP_PX - 1) / mcl::LAP_PX) * ((max_h + mcl::LAP_RY - 1) / mcl::LAP_RY);
        dim3 grid((unsigned)((threads + 255) / 256), n);
        ProfScope ps(c, PROF_LAP, st);
        mcl::laplacian_sums_kernel<<<grid, 256, 0, st>>>(c->lap_pix.as<uint8_t>(), (const int64_t*)mb, (const int*)(mb + o_p),
                                                         (const int*)(mb + o_h), (const int*)(mb + o_w),
                                                         (unsigned long long*)(mb + o_s));
    }
    CU(cudaGetLastError());
    c->launches++;
    mcl::laplacian_finish_kernel<<<(n + 3) / 4, 128, 0, st>>>(c->lap_pix.as<uint8_t>(), (const int64_t*)mb, (const int*)(mb + o_p),
                                                              (const int*)(mb + o_h), (const int*)(mb + o_w), n,
                                                              (long long*)(mb + o_s), (double*)(mb + o_v));
    CU(cudaGetLastError());
    if (var_out) {
        // numpy's two-pass variance, addition for addition (lap_var.cuh); images grouped by pixel count (one table each)
        std::map<int64_t, std::vector<int>> groups;
        for (int i = 0; i < n; ++i) {
            if ((int64_t)h[i] * w[i] >= ((int64_t)1 << 31)) return fail(MCL_EINVAL, "image %d: too many pixels", i);
            groups[(int64_t)h[i] * w[i]].push_back(i);
        }
        std::vector<int> list;
        for (auto& g : groups) list.insert(list.end(), g.second.begin(), g.second.end());
        CU(c->lap_list.ensure((size_t)n * 4));
        CU(cudaMemcpyAsync(c->lap_list.p, list.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
        size_t need = 0;
        for (auto& g : groups) {
            const LapPlan* pl = nullptr;
            int r = lap_plan_get(c, g.first, &pl);
            if (r) return r;
            need = std::max(need, (size_t)(pl->n_leaves + pl->n_nodes) * g.second.size());
        }
        CU(c->lap_vals.ensure(need * 8));
        int first = 0;
        for (auto& g : groups) {
            const LapPlan* pl = &c->lap_plans[g.first];
            const int cnt = (int)g.second.size(), nl = pl->n_leaves, nn = pl->n_nodes, nv = nl + nn;
            const int* t = pl->d_tab;
            const int* lst = c->lap_list.as<int>() + first;
            const int bx = std::max(1, std::min((nl + 31) / 32, std::max(1, 8 * c->sm_count / cnt)));
            // widths of the group (images of one pixel count may still differ in shape): the thread-per-leaf kernel needs
            // W % 8 == 0 and its staged rows (128 leaves x <= 128 pixels + 3 rows) within the shared-memory opt-in
            bool fast = true;
            int wmax = 0;
            for (int i : g.second) { fast = fast && (w[i] % 8 == 0); wmax = std::max(wmax, w[i]); }
            const size_t smem = (size_t)mcl::LAPV_LEAVES_PER_CTA * 128 + 4 * (size_t)wmax;
            fast = fast && smem <= (size_t)c->prop.sharedMemPerBlockOptin;
            if (fast && smem > (size_t)c->lap_smem_optin) {
                CU(cudaFuncSetAttribute(mcl::lap_leaf8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                c->lap_smem_optin = (int)smem;
            }
            {
                ProfScope ps(c, PROF_LAP, st);
                if (fast)
                    mcl::lap_leaf8_kernel<<<dim3((nl + mcl::LAPV_LEAVES_PER_CTA - 1) / mcl::LAPV_LEAVES_PER_CTA, cnt), mcl::LAPV_LEAVES_PER_CTA, smem, st>>>(
                        c->lap_pix.as<uint8_t>(), (const int64_t*)mb, (const int*)(mb + o_p), (const int*)(mb + o_h), (const int*)(mb + o_w),
                        lst, (const long long*)(mb + o_s), t, t + nl, nl, nv, c->lap_vals.as<double>());
                else
                    mcl::lap_leaf_kernel<<<dim3(bx, cnt), 256, 0, st>>>(c->lap_pix.as<uint8_t>(), (const int64_t*)mb, (const int*)(mb + o_p),
                                                                        (const int*)(mb + o_h), (const int*)(mb + o_w), lst,
                                                                        (const long long*)(mb + o_s), t, t + nl, nl, nv, c->lap_vals.as<double>());
            }
            c->launches++;
            mcl::lap_tree_kernel<<<cnt, 256, 0, st>>>(t + 2 * nl, t + 2 * nl + nn, t + 2 * nl + 2 * nn, pl->n_levels, nl, nv, lst,
                                                      (const int*)(mb + o_h), (const int*)(mb + o_w), c->lap_vals.as<double>(),
                                                      (double*)(mb + o_v));
            CU(cudaGetLastError());
            first += cnt;
        }
        CU(cudaMemcpyAsync(var_out, mb + o_v, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    }
    if (sums_out) CU(cudaMemcpyAsync(sums_out, mb + o_s, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return MCL_OK;
}
}  // namespace

extern "C" int mcl_laplacian_var(mcl_ctx* c, int n, const uint8_t* const* imgs, const int* h, const int* w, const int* stride,
                                 double* out) {
    if (n > 0 && !out) return fail(MCL_EINVAL, "mcl_laplacian_var: out is NULL");
    return laplacian_common(c, n, imgs, h, w, stride, nullptr, out);
}
extern "C" int mcl_laplacian_var_packed(mcl_ctx* c, int n, const uint8_t* packed, const int* h, const int* w, double* out) {
    if (n > 0 && (!out || !packed)) return fail(MCL_EINVAL, "mcl_laplacian_var_packed: NULL argument");
    return laplacian_common(c, n, nullptr, h, w, nullptr, nullptr, out, packed);
}
extern "C" int mcl_laplacian_sums(mcl_ctx* c, int n, const uint8_t* const* imgs, const int* h, const int* w, const int* stride,
                                  int64_t* sums) {
    if (n > 0 && !sums) return fail(MCL_EINVAL, "mcl_laplacian_sums: sums is NULL");
    return laplacian_common(c, n, imgs, h, w, stride, sums, nullptr);
}

namespace {
int filter_step_common(mcl_ctx* c, int L, int I, int stages, char cmd, double blur, int blur_type, double* prev,
                       const int8_t* prev_type, const double* cur, const int8_t* cur_type, double* out, int8_t* out_type,
                       int* best) {
    if (!c || !prev || !cur || !out || !best) return fail(MCL_EINVAL, "mcl_filter_step: NULL argument");
    if (L <= 0 || I <= 0) return fail(MCL_EINVAL, "mcl_filter_step: L and I must be positive");
    if (stages <= 0 || stages > MCL_F_ALL) return fail(MCL_EINVAL, "mcl_filter_step: bad stages mask");
    if (blur_type < 0 || blur_type > 2) return fail(MCL_EINVAL, "mcl_filter_step: bad blur type");
    for (int l = 0; l < 2 * L; ++l)
        if ((prev_type && (prev_type[l] < 0 || prev_type[l] > 2)) || (cur_type && (cur_type[l] < 0 || cur_type[l] > 2)))
            return fail(MCL_EINVAL, "mcl_filter_step: row %d: type tag must be MCL_T_PY, MCL_T_F64 or MCL_T_F32", l / 2);
    CU(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t n = (size_t)L * (1 + I);
    // layout: prev | cur | out | scratch | factor (I) | best (2 ints) | prev types | cur types | out types (2L bytes each)
    const size_t total = (4 * n + I) * 8 + 16 + 6 * (size_t)L;
    CU(c->filt.ensure(total));
    double* d = c->filt.as<double>();
    double *d_prev = d, *d_cur = d + n, *d_out = d + 2 * n, *d_scr = d + 3 * n, *d_fac = d + 4 * n;
    int* d_best = reinterpret_cast<int*>(d_fac + I);
    signed char* d_pt = reinterpret_cast<signed char*>(d_best + 4);
    signed char *d_ct = d_pt + 2 * L, *d_ot = d_ct + 2 * L;
    std::vector<double> fac(I);
    const int step_deg = c->opt[6] > 0 ? c->opt[6] : 15;   // option 6: degrees between neighbouring map images (reference: 15)
    for (int k = 0; k < I; ++k) fac[k] = 0.05 * std::fabs(std::sin((double)(k * step_deg * 180) / M_PI));  // analyze.py:331 (sic)
    CU(cudaMemcpyAsync(d_prev, prev, n * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_cur, cur, n * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_fac, fac.data(), (size_t)I * 8, cudaMemcpyHostToDevice, st));
    if (prev_type) CU(cudaMemcpyAsync(d_pt, prev_type, (size_t)2 * L, cudaMemcpyHostToDevice, st));
    if (cur_type) CU(cudaMemcpyAsync(d_ct, cur_type, (size_t)2 * L, cudaMemcpyHostToDevice, st));
    c->launches++;
    mcl::filter_step_kernel<<<1, 256, 0, st>>>(L, I, step_deg, stages, (int)cmd, blur, blur_type, d_prev, prev_type ? d_pt : nullptr, d_cur,
                                               cur_type ? d_ct : nullptr, d_fac, d_out, out_type ? d_ot : nullptr, d_best, d_scr);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(prev, d_prev, n * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out, d_out, n * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(best, d_best, 8, cudaMemcpyDeviceToHost, st));
    if (out_type) {
        if (stages & 6) CU(cudaMemcpyAsync(out_type, d_ot, (size_t)2 * L, cudaMemcpyDeviceToHost, st));
        else std::memset(out_type, 0, (size_t)2 * L);
    }
    CU(cudaStreamSynchronize(st));
    return MCL_OK;
}
}  // namespace

extern "C" int mcl_filter_step(mcl_ctx* c, int L, int I, int stages, char cmd, double blur, double* prev, const double* cur,
                               double* out, int* best) {
    return filter_step_common(c, L, I, stages, cmd, blur, 0, prev, nullptr, cur, nullptr, out, nullptr, best);
}
extern "C" int mcl_filter_step_typed(mcl_ctx* c, int L, int I, int stages, char cmd, double blur, int blur_type, double* prev,
                                     const int8_t* prev_type, const double* cur, const int8_t* cur_type, double* out,
                                     int8_t* out_type, int* best) {
    return filter_step_common(c, L, I, stages, cmd, blur, blur_type, prev, prev_type, cur, cur_type, out, out_type, best);
}

extern "C" int mcl_chi2(mcl_ctx* c, int n_q, int n_map, int bins, const float* q_hist, const float* map_hist, float eps,
                        float* out) {
    if (!c) return fail(MCL_EINVAL, "mcl_chi2: NULL context");
    if (n_q < 0 || n_map < 0 || bins <= 0) return fail(MCL_EINVAL, "mcl_chi2: negative size");
    if (n_q == 0 || n_map == 0) return MCL_OK;
    if (!q_hist || !map_hist || !out) return fail(MCL_EINVAL, "mcl_chi2: NULL argument");
    if (n_q > 65535) return fail(MCL_EINVAL, "mcl_chi2: at most 65535 query histograms per call");
    if ((int64_t)bins > (1 << 24)) return fail(MCL_EINVAL, "mcl_chi2: too many bins");
    CU(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t qb = (size_t)n_q * bins * 4, mb = (size_t)n_map * bins * 4, ob = (size_t)n_q * n_map * 4;
    const size_t o_m = (qb + 255) & ~(size_t)255, o_o = (o_m + mb + 255) & ~(size_t)255;
    CU(c->chi2.ensure(o_o + ob));
    uint8_t* d = c->chi2.as<uint8_t>();
    CU(cudaMemcpyAsync(d, q_hist, qb, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d + o_m, map_hist, mb, cudaMemcpyHostToDevice, st));
    {
        ProfScope ps(c, PROF_CHI2, st);
        if (bins == 512) {
            // n_q rows of CTAs; enough CTAs per row to fill the SMs a few times over when n_q is small
            const int gx = grid_for(n_map, mcl::CHI2_WARPS, std::max(1, 4 * c->sm_count / n_q));
            mcl::chi2_512_kernel<<<dim3(gx, n_q), mcl::CHI2_WARPS * 32, 0, st>>>((const float*)d, (const float*)(d + o_m), n_q, n_map,
                                                                                eps, (float*)(d + o_o));
        } else {
            const long long pairs = (long long)n_q * n_map;
            mcl::chi2_generic_kernel<<<(unsigned)((pairs + 127) / 128), 128, 0, st>>>((const float*)d, (const float*)(d + o_m), n_q,
                                                                                     n_map, bins, eps, (float*)(d + o_o));
        }
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, d + o_o, ob, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return MCL_OK;
}

// ================================================================================================
// peak microbenchmarks (roofline denominators that MEASURED_PEAKS.json does not carry: int8 tensor pipe, popc pipe)
// ================================================================================================
extern "C" int mcl_microbench(mcl_ctx* c, int which, int iters, double* out /* [0]=ops/s, [1]=ms */) {
    if (!c || !out) return fail(MCL_EINVAL, "mcl_microbench: NULL argument");
    CU(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a));
    CU(cudaEventCreate(&b));
    double ops = 0;
    CU(c->stage.ensure(1 << 20));
    if (which == 0) {  // tcgen05 kind::i8, 128x128x32 MMAs back to back from shared memory, one CTA per SM
        const size_t smem = 2 * 16384 + 1024 + 64;
        CU(cudaFuncSetAttribute(mcl::i8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mcl::i8_peak_kernel<<<c->sm_count, 128, smem, st>>>(iters / 8 + 1, c->stage.as<int>());  // warm-up
        CU(cudaEventRecord(a, st));
        mcl::i8_peak_kernel<<<c->sm_count, 128, smem, st>>>(iters, c->stage.as<int>());
        CU(cudaEventRecord(b, st));
        ops = 2.0 * 128 * 128 * 32 * 16.0 * (double)iters * c->sm_count;
    } else if (which == 1) {  // popc.b32 + xor issue rate
        mcl::popc_peak_kernel<<<c->sm_count * 8, 256, 0, st>>>(iters / 8 + 1, c->stage.as<int>());
        CU(cudaEventRecord(a, st));
        mcl::popc_peak_kernel<<<c->sm_count * 8, 256, 0, st>>>(iters, c->stage.as<int>());
        CU(cudaEventRecord(b, st));
        ops = 16.0 * (double)iters * 256.0 * c->sm_count * 8;  // popc32 results
    } else if (which == 2) {  // fp32 FADD + FFMA pairs (SURF's direct-difference form): issue rate of the FMA pipe
        mcl::ffma_peak_kernel<<<c->sm_count * 8, 256, 0, st>>>(iters / 8 + 1, c->stage.as<int>());
        CU(cudaEventRecord(a, st));
        mcl::ffma_peak_kernel<<<c->sm_count * 8, 256, 0, st>>>(iters, c->stage.as<int>());
        CU(cudaEventRecord(b, st));
        ops = 16.0 * (double)iters * 256.0 * c->sm_count * 8;  // (FADD, FFMA) element pairs
    } else if (which >= 50 && which < 60) {  // development probe: integer-pipe issue rates, which = 50 + OP
        void (*k)(int, int*) = nullptr;
        switch (which - 50) {
            case 0: k = mcl::alu_probe_kernel<0>; break;
            case 1: k = mcl::alu_probe_kernel<1>; break;
            case 2: k = mcl::alu_probe_kernel<2>; break;
            case 3: k = mcl::alu_probe_kernel<3>; break;
            case 4: k = mcl::alu_probe_kernel<4>; break;
            case 5: k = mcl::alu_probe_kernel<5>; break;
            case 6: k = mcl::alu_probe_kernel<6>; break;
            default: return fail(MCL_EINVAL, "mcl_microbench: unknown probe variant");
        }
        k<<<c->sm_count * 8, 256, 0, st>>>(iters / 8 + 1, c->stage.as<int>());
        CU(cudaEventRecord(a, st));
        k<<<c->sm_count * 8, 256, 0, st>>>(iters, c->stage.as<int>());
        CU(cudaEventRecord(b, st));
        ops = 16.0 * (double)iters * 256.0 * c->sm_count * 8;
    } else if (which >= 100) {  // development probe: which = 100 + n/64 (1,2,4) + 10*ts + 100*ext
        const int w = which - 100;
        const int n = (w % 10) * 64, ts = (w / 10) % 10, ext = (w / 100) % 10;
        const size_t smem = 16384 + 32768 + 8192 + 1024 + 64;
        void (*k)(int, int*) = nullptr;
#define PICK(N_, TS_, EXT_) if (n == N_ && ts == TS_ && ext == EXT_) k = mcl::mma_probe_kernel<N_, TS_, EXT_>;
        PICK(64, 0, 0) PICK(64, 1, 0) PICK(64, 0, 1) PICK(64, 1, 1)
        PICK(128, 0, 0) PICK(128, 1, 0) PICK(128, 0, 1) PICK(128, 1, 1)
        PICK(256, 0, 0) PICK(256, 1, 0) PICK(256, 0, 1) PICK(256, 1, 1)
#undef PICK
        if (!k) return fail(MCL_EINVAL, "mcl_microbench: unknown probe variant");
        CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<c->sm_count, 128, smem, st>>>(iters / 8 + 1, c->stage.as<int>());
        CU(cudaEventRecord(a, st));
        k<<<c->sm_count, 128, smem, st>>>(iters, c->stage.as<int>());
        CU(cudaEventRecord(b, st));
        ops = 2.0 * 128 * n * 32 * 15.0 * (double)iters * c->sm_count;
    } else {
        return fail(MCL_EINVAL, "mcl_microbench: which must be 0 (int8 tensor), 1 (popc) or 2 (fp32 difference-square pairs)");
    }
    c->launches += 2;
    CU(cudaGetLastError());
    CU(cudaEventSynchronize(b));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    out[0] = ops / (ms * 1e-3);
    out[1] = ms;
    return MCL_OK;
}
