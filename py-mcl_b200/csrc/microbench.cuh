// Peak microbenchmarks for the two pipes that bound the hot path and that MEASURED_PEAKS.json does not carry.
#pragma once
#include "common.cuh"

namespace mcl {

// tcgen05.mma kind::i8 (u8 x u8 -> s32), M=128 N=128 K=32 per instruction, operands resident in shared memory.
// One CTA per SM, one issuing thread; 16 MMAs per iteration into two accumulators (256 TMEM columns).
__global__ void __launch_bounds__(128, 1) i8_peak_kernel(int iters, int* __restrict__ sink) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a = base;
    uint8_t* b = base + 16384;
    uint64_t* bar = reinterpret_cast<uint64_t*>(base + 32768);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base + 32768 + 16);
    for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x01010101u * (i & 3);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    // make the generic-proxy smem writes visible to the async proxy (tensor core reads)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = umma_idesc(2, 0, 0, 128, 128);
        const uint64_t da = umma_desc_sw128(smem_u32(a));
        const uint64_t db = umma_desc_sw128(smem_u32(b));
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_i8(tmem + (r & 1) * 128, da + 2 * k, db + 2 * k, idesc, (it | r | k) != 0);
            }
        }
        tc_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    tc_fence_after();
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16), v);
    tmem_ld_wait(v);
    if (v[0] == 0xdeadbeefu) sink[blockIdx.x] = (int)v[1];
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<256>(tmem);
}

// tcgen05 shape probe: N (64/128/256), A from shared memory (TS=0) or tensor memory (TS=1); with EXT every 5th MMA
// reads a no-swizzle 32-byte-row B operand (the K-extension step of sift_tc2).  15 MMAs per iteration, fully unrolled,
// issued by the elected lane of a converged warp (uniform operands).  The host times it.
template <int N, int TS, int EXT>
__global__ void __launch_bounds__(128, 1) mma_probe_kernel(int iters, int* __restrict__ sink) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a = base;                  // 128 x 128 B
    uint8_t* b = base + 16384;          // 256 x 128 B
    uint8_t* e = base + 16384 + 32768;  // 256 x 32 B no-swizzle
    uint64_t* bar = reinterpret_cast<uint64_t*>(base + 16384 + 32768 + 8192);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
    for (int i = threadIdx.x; i < (16384 + 32768 + 8192) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x01010101u * (i & 3);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int warp = warp_id_uniform();
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp == 0) {
        const uint32_t leader = (threadIdx.x & 31) == 0;
        constexpr uint32_t idesc = umma_idesc(2, 0, 0, 128, N);
        const uint64_t da = umma_desc_sw128(smem_u32(a));
        const uint64_t db = umma_desc_sw128(smem_u32(b));
        const uint64_t de = umma_desc_noswizzle(smem_u32(e), 128, 256);
        constexpr int NACC = N >= 256 ? 1 : (N == 128 ? 2 : 3);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 15; ++r) {
                const int k = r % 5;
                const uint32_t d = tmem + 128 + ((r / 5) % NACC) * N;
                const uint64_t bd = (EXT && k == 4) ? de : db + 2 * (k & 3);
                if (TS) umma_i8_ts_if(leader, d, tmem + 8 * k, bd, idesc, 1);
                else umma_i8_if(leader, d, da + 2 * (k & 3), bd, idesc, 1);
            }
        }
        tc_commit_if(leader, bar);
        mbar_wait(bar, 0);
    }
    __syncthreads();
    tc_fence_after();
    uint32_t v[16];
    tmem_ld16(tmem + 128 + ((uint32_t)((threadIdx.x >> 5) * 32) << 16), v);
    tmem_ld_wait(v);
    if (v[0] == 0xdeadbeefu) sink[blockIdx.x] = (int)v[1];
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// integer-pipe issue-rate probe: 16 independent chains per thread, one op of the selected kind per chain per iteration.
// OP: 0 = max3 (VIMNMX3), 1 = max (VIMNMX), 2 = add3 (IADD3), 3 = mad (IMAD), 4 = fmaxf (FMNMX), 5 = max3 + mad interleaved,
// 6 = 16-bit packed max (VIMNMX.S16x2 via __vmaxs2)
template <int OP>
__global__ void __launch_bounds__(256) alu_probe_kernel(int iters, int* __restrict__ sink) {
    int x[16], y[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        x[i] = threadIdx.x * 2654435 + i * 40503 + blockIdx.x;
        y[i] = x[i] ^ 0x5bd1e995;
    }
    int z = threadIdx.x ^ 0x9e3779b9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (OP == 0) x[i] = max3(x[i], y[i], z);
            if (OP == 1) x[i] = max(x[i], y[i] + 0 * z);
            if (OP == 2) x[i] = x[i] + y[i] + z;
            if (OP == 3) x[i] = x[i] * z + y[i];
            if (OP == 4) x[i] = __float_as_int(fmaxf(__int_as_float(x[i]), __int_as_float(y[i])));
            if (OP == 5) { x[i] = max3(x[i], y[i], z); y[i] = y[i] * z + x[(i + 1) & 15]; }
            if (OP == 6) x[i] = (int)__vmaxs2((unsigned)x[i], (unsigned)y[i]);
        }
        z += 0x7f4a7c15;
        if (OP == 1 || OP == 4 || OP == 6) {
#pragma unroll
            for (int i = 0; i < 16; ++i) y[i] ^= z;   // keeps the max from being hoisted; one extra LOP3 per op
        }
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i] ^ y[i];
    if (s == 0x7fffffff) sink[0] = s;
}

// XOR + popc.b32: 16 independent results per iteration per thread
__global__ void __launch_bounds__(256) popc_peak_kernel(int iters, int* __restrict__ sink) {
    unsigned x[16];
    int acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        x[i] = threadIdx.x * 2654435761u + i * 40503u + blockIdx.x;
        acc[i] = 0;
    }
    unsigned y = threadIdx.x ^ 0x9e3779b9u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] += __popc(x[i] ^ y);
        y += 0x7f4a7c15u;
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 0x7fffffff) sink[0] = s;
}

// fp32 issue rate for the direct-difference form SURF uses: one FADD (a - b) and one FFMA (d * d + s) per element, 16
// independent chains per thread.  out: element pairs (FADD + FFMA) per second.
__global__ void __launch_bounds__(256) ffma_peak_kernel(int iters, int* __restrict__ sink) {
    float a[16], s[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        a[i] = (float)(threadIdx.x + i) * 1e-3f;
        s[i] = 0.f;
    }
    float b = (float)blockIdx.x * 1e-4f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float d = a[i] - b;
            s[i] = fmaf(d, d, s[i]);
        }
        b += 1e-6f;
    }
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += s[i];
    if (t == 123.456f) sink[0] = 1;
}

}  // namespace mcl
