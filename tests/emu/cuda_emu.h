// TEST INFRASTRUCTURE ONLY -- a minimal CUDA execution-model emulator for g++.
//
// The build container has nvcc but no GPU.  To validate kernel *logic* (tile
// indexing, halos, shared-memory hand-offs, warp/block reductions, adjoint
// formulas) before spending GPU minutes, tests/emu/build_emu.sh compiles the
// very same csrc/*.cu sources with g++ against this header.  Each CUDA block is
// run by blockDim OS threads; __syncthreads() is a std::barrier, warp shuffles
// exchange through a per-warp slot array, atomics are std::atomic_ref.
//
// The resulting libtcsfm_emu.so is loaded only by tests/ (never by the package:
// tcsfm_b200/_lib.py loads libtcsfm_b200.so and nothing else, and the operators
// refuse CPU tensors).  It is not a fallback and it is never timed.
#pragma once
#ifndef TCSFM_HOST_EMU
#error "cuda_emu.h is only for -DTCSFM_HOST_EMU builds"
#endif

#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }

typedef int cudaError_t;
typedef void* cudaStream_t;
#define cudaSuccess 0

namespace emu {
struct BlockCtx {
    std::unique_ptr<std::barrier<>> block_barrier;
    std::vector<std::unique_ptr<std::barrier<>>> warp_barrier;
    std::vector<uint32_t> slots;   // [warps][32] shuffle exchange
    std::vector<char> dyn_smem;
};
inline BlockCtx*& ctx() { static BlockCtx* c = nullptr; return c; }
}  // namespace emu

extern thread_local dim3 threadIdx;
extern thread_local dim3 blockIdx;
extern thread_local dim3 blockDim;
extern thread_local dim3 gridDim;
#ifdef TCSFM_EMU_DEFINE_GLOBALS
thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
#endif

static inline unsigned emu_linear_tid() {
    return threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
}
static inline void __syncthreads() { emu::ctx()->block_barrier->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) {
    emu::ctx()->warp_barrier[emu_linear_tid() / 32]->arrive_and_wait();
}
template <typename T>
static inline T emu_shfl(T v, unsigned src_lane) {
    static_assert(sizeof(T) == 4, "32-bit shuffles only");
    auto* c = emu::ctx();
    unsigned tid = emu_linear_tid(), w = tid / 32, l = tid % 32;
    uint32_t bits;
    std::memcpy(&bits, &v, 4);
    c->slots[w * 32 + l] = bits;
    c->warp_barrier[w]->arrive_and_wait();
    uint32_t got = c->slots[w * 32 + (src_lane & 31)];
    c->warp_barrier[w]->arrive_and_wait();
    T out;
    std::memcpy(&out, &got, 4);
    return out;
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_shfl(v, (emu_linear_tid() % 32) ^ (unsigned)m); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int d) {
    unsigned l = emu_linear_tid() % 32;
    return emu_shfl(v, l + d < 32 ? l + d : l);
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) { return emu_shfl(v, (unsigned)src); }

static inline float atomicAdd(float* addr, float v) {
    std::atomic_ref<float> r(*addr);
    return r.fetch_add(v, std::memory_order_relaxed);
}
static inline int atomicAdd(int* addr, int v) {
    std::atomic_ref<int> r(*addr);
    return r.fetch_add(v, std::memory_order_relaxed);
}
static inline unsigned atomicAdd(unsigned* addr, unsigned v) {
    std::atomic_ref<unsigned> r(*addr);
    return r.fetch_add(v, std::memory_order_relaxed);
}
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

// IEEE single ops; build with -ffp-contract=off so plain * and + never fuse.
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fdividef(float a, float b) { return a / b; }
// sm_100 packed fp32x2 arithmetic (crt/sm_100_rt.h): IEEE per lane
static inline float2 __fadd2_rn(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
static inline float2 __fmul2_rn(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)}; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline int __float2int_rd(float x) {
    float f = std::floor(x);
    if (!(f == f)) return 0;                     // cvt of NaN is 0 on the GPU
    if (f >= 2147483648.0f) return 2147483647;   // saturating like cvt.rmi.s32.f32
    if (f <= -2147483648.0f) return (-2147483647 - 1);
    return (int)f;
}
static inline float __int2float_rn(int x) { return (float)x; }
static inline float __int_as_float(int x) { float f; std::memcpy(&f, &x, 4); return f; }
static inline int __float_as_int(float f) { int x; std::memcpy(&x, &f, 4); return x; }

// cp.async (cuda_pipeline.h primitives): a plain copy with trailing zero fill on the host
static inline void __pipeline_memcpy_async(void* dst, const void* src, size_t size, size_t zfill = 0) {
    std::memcpy(dst, src, size - zfill);
    std::memset((char*)dst + (size - zfill), 0, zfill);
}
static inline void __pipeline_commit() {}
static inline void __pipeline_wait_prior(int) {}

static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }

namespace emu {
// Runs `body` once per (block, thread).  All blockDim threads advance through the
// grid in lock-step, one block at a time (statics standing in for __shared__ are
// therefore private to the running block).
template <typename F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, F body) {
    unsigned nthreads = block.x * block.y * block.z;
    if (nthreads % 32 != 0) { std::fprintf(stderr, "emu: block size must be a multiple of 32\n"); std::abort(); }
    BlockCtx c;
    c.block_barrier = std::make_unique<std::barrier<>>(nthreads);
    for (unsigned w = 0; w < nthreads / 32; ++w) c.warp_barrier.push_back(std::make_unique<std::barrier<>>(32));
    c.slots.assign(nthreads, 0);
    c.dyn_smem.assign(smem_bytes + 64, 0);
    ctx() = &c;
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nthreads; ++t) {
        pool.emplace_back([&, t]() {
            blockDim = block; gridDim = grid;
            threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            for (unsigned bz = 0; bz < grid.z; ++bz)
                for (unsigned by = 0; by < grid.y; ++by)
                    for (unsigned bx = 0; bx < grid.x; ++bx) {
                        blockIdx = dim3(bx, by, bz);
                        body();
                        c.block_barrier->arrive_and_wait();
                    }
        });
    }
    for (auto& th : pool) th.join();
    ctx() = nullptr;
}
inline void* dyn_smem_base() {
    uintptr_t p = (uintptr_t)ctx()->dyn_smem.data();
    return (void*)((p + 63) & ~(uintptr_t)63);
}
}  // namespace emu
