// Shared launch / reduction helpers for the tcsfm sm_100a kernels.
//
// The same sources are also compiled by g++ against tests/emu/cuda_emu.h
// (-DTCSFM_HOST_EMU) so that kernel logic can be checked without a GPU; that
// build is test infrastructure only and is never loaded by the package.
#pragma once

#ifdef TCSFM_HOST_EMU
#include "cuda_emu.h"
#define TCSFM_SHARED static
#define TCSFM_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::dyn_smem_base())
#define TCSFM_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })
#else
#include <cuda_runtime.h>
#include <cuda_pipeline.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>
#define TCSFM_SHARED __shared__
#define TCSFM_DYN_SMEM(type, name) extern __shared__ __align__(16) unsigned char name##_raw_[]; \
    type* name = reinterpret_cast<type*>(name##_raw_)
#define TCSFM_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
#endif

#include "../../include/tcsfm.h"

namespace tcsfm {

// ---- error reporting (per-thread last error string, C ABI: tcsfm_last_error) ----
void set_error(const char* fmt, ...);
int check_launch(const char* what);

constexpr int kWarp = 32;

// Makes a per-thread base pointer opaque to the compiler so that it stays in registers:
// without this ptxas re-derives `base + b * stride` (a 64-bit multiply-add chain from the
// kernel parameters) at every use to save two registers.
template <class T>
__device__ __forceinline__ T* pin_pointer(T* p) {
#ifndef TCSFM_HOST_EMU
    asm volatile("" : "+l"(p));
#endif
    return p;
}

__device__ __forceinline__ float warp_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// Store through a pinned pointer: the asm above hides the address space from the compiler,
// which would otherwise emit a generic ST instead of STG.
__device__ __forceinline__ void store_global(float* p, float v) {
#ifdef TCSFM_HOST_EMU
    *p = v;
#else
    asm volatile("st.global.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
#endif
}

// Workspace stores: written once, read once by the backward much later -- the evict-first hint keeps them from
// pushing the images / depths (which the other pair groups of the frame are about to gather) out of L2.
__device__ __forceinline__ void store_streaming(float* p, float v) {
#if defined(TCSFM_HOST_EMU)
    *p = v;
#elif defined(TCSFM_WS_NO_STREAMING)
    asm volatile("st.global.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
#else
    asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
#endif
}
__device__ __forceinline__ void store_streaming2(float* p, float2 v) {
#if defined(TCSFM_HOST_EMU)
    p[0] = v.x; p[1] = v.y;
#elif defined(TCSFM_WS_NO_STREAMING)
    asm volatile("st.global.v2.f32 [%0], {%1, %2};" :: "l"(p), "f"(v.x), "f"(v.y) : "memory");
#else
    asm volatile("st.global.cs.v2.f32 [%0], {%1, %2};" :: "l"(p), "f"(v.x), "f"(v.y) : "memory");
#endif
}

// One 4-byte cp.async global -> shared; `live == false` writes a zero instead of reading
// `src` (the ignore-src form: a single LDGSTS with a predicate operand, where the
// cuda_pipeline.h helper with a run-time zfill emits two predicated copies).
__device__ __forceinline__ void async_copy4(float* smem_dst, const float* src, bool live) {
#ifdef TCSFM_HOST_EMU
    *smem_dst = live ? *src : 0.f;
#else
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %2, 0;\n\t"
                 "cp.async.ca.shared.global [%0], [%1], 4, p;\n\t}"
                 :: "r"(dst), "l"(src), "r"((unsigned)live) : "memory");
#endif
}

// Two consecutive floats (both addresses 8-byte aligned).
__device__ __forceinline__ void async_copy8(float* smem_dst, const float* src, bool live) {
#ifdef TCSFM_HOST_EMU
    for (int i = 0; i < 2; ++i) smem_dst[i] = live ? src[i] : 0.f;
#else
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %2, 0;\n\t"
                 "cp.async.ca.shared.global [%0], [%1], 8, p;\n\t}"
                 :: "r"(dst), "l"(src), "r"((unsigned)live) : "memory");
#endif
}

// The same for four consecutive floats (both addresses 16-byte aligned); .cg: streamed once, L2 only.
__device__ __forceinline__ void async_copy16(float* smem_dst, const float* src, bool live) {
#ifdef TCSFM_HOST_EMU
    for (int i = 0; i < 4; ++i) smem_dst[i] = live ? src[i] : 0.f;
#else
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %2, 0;\n\t"
                 "cp.async.cg.shared.global [%0], [%1], 16, p;\n\t}"
                 :: "r"(dst), "l"(src), "r"((unsigned)live) : "memory");
#endif
}

// Warp-wide sums of N <= 16 per-thread values with 16 shuffles instead of 5 N: at every butterfly stage a lane keeps
// one half of its values and hands the other half to its partner, so the number of live values halves with the
// distance.  On return v[0] of lane L holds the warp total of value index ((L >> 1) & 15) (both lanes of a pair hold it).
template <int N>
__device__ __forceinline__ float warp_sum_transposed(const float (&part)[N], int lane) {
    static_assert(N <= 16, "warp_sum_transposed handles at most 16 values");
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = i < N ? part[i] : 0.f;
#pragma unroll
    for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
        const bool upper = (lane & bit) != 0;           // this lane keeps the upper half of its live values
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float keep = upper ? v[half + i] : v[i];
            const float send = upper ? v[i] : v[half + i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// Block-wide sum of N per-thread partials -> one atomicAdd per value per block.
// `red` is shared scratch of at least N * (threads/32) floats.  All threads of
// the block must call this (it contains __syncthreads()).
template <int N>
__device__ __forceinline__ void block_atomic_accumulate(const float (&part)[N], float* red, float* dst,
                                                        int tid, int nthreads) {
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    if (N > 4) {
        const float s = warp_sum_transposed<N>(part, lane);
        const int idx = (lane >> 1) & 15;
        if (!(lane & 1) && idx < N) red[idx * nwarps + warp] = s;
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            float s = warp_sum(part[i]);
            if (lane == 0) red[i * nwarps + warp] = s;
        }
    }
    __syncthreads();
    if (tid < N) {
        float s = 0.f;
        for (int w = 0; w < nwarps; ++w) s += red[tid * nwarps + w];
        if (s != 0.f) atomicAdd(dst + tid, s);
    }
    __syncthreads();
}

}  // namespace tcsfm
