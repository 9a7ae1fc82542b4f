// Stand-alone probe of the TMA staging used by pair_bwd_kernel<F, true>: loads a 68 x 18 x 9 x 1 box of a
// [B][10][H][W] fp32 tensor (zero fill outside) and checks it against direct loads.  nvcc -arch=sm_100a, run on the GPU.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct Launch { int x0, y0, b, H, W; alignas(64) CUtensorMap map[2]; };
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ Launch L, float* out, int mode, const CUtensorMap* gmap, int boxw) {
    extern __shared__ __align__(128) unsigned char raw[];
    float* cs = reinterpret_cast<float*>(raw);
    void* bar = cs + 68 * 18 * 9;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(1u) : "memory");
        if (!(mode & 4)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(bar)), "r"((unsigned)boxw * 18u * 9u * 4u) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     :: "r"(smem_addr(cs)), "l"((mode & 1) ? gmap : &L.map[1]), "r"(smem_addr(bar)), "r"(L.x0 - 3), "r"(L.y0 - 1), "r"(0), "r"(L.b) : "memory");
    }
    unsigned spins = 0;
    for (;; ++spins) {
        unsigned done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_addr(bar)), "r"(0u) : "memory");
        if (done) break;
        if (spins > (1u << 22)) { if (threadIdx.x == 0) out[0] = -12345.f; return; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 68 * 18 * 9; i += blockDim.x) out[1 + i] = cs[i];
    if (threadIdx.x == 0) out[0] = (float)spins;
}
int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int boxw = (mode & 2) ? 64 : 68;
    const int B = 2, H = 64, W = 96, P = 10;
    std::vector<float> h((size_t)B * P * H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 9973) + 1.0f;
    float *d, *o;
    cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&o, (1 + 68 * 18 * 9) * 4); cudaMemset(o, 0, (1 + 68 * 18 * 9) * 4);
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    printf("entry point: %s q=%d ptr=%p\n", cudaGetErrorString(e), (int)q, fp);
    if (!fp) return 1;
    Launch L; memset(&L, 0, sizeof(L));
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)P, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)P * H * W * 4};
    const cuuint32_t box[4] = {(cuuint32_t)boxw, 18, 9, 1}, estr[4] = {1, 1, 1, 1};
    CUresult r = ((EncodeTiledFn)fp)(&L.map[1], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("mode %d boxw %d encode: %d  sizeof(Launch)=%zu alignof=%zu\n", mode, boxw, (int)r, sizeof(Launch), alignof(Launch));
    CUtensorMap* gmap; cudaMalloc(&gmap, sizeof(CUtensorMap)); cudaMemcpy(gmap, &L.map[1], sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    int bad_total = 0;
    for (int t = 0; t < 3; ++t) {
        L.x0 = t == 1 ? 64 : 0; L.y0 = t == 2 ? 48 : 0; L.b = t % 2; L.H = H; L.W = W;
        const size_t smem = 68 * 18 * 9 * 4 + 16;
        probe<<<1, 256, smem>>>(L, o, mode, gmap, boxw);
        e = cudaDeviceSynchronize();
        printf("case %d: launch/sync: %s\n", t, cudaGetErrorString(e));
        if (e != cudaSuccess) return 2;
        std::vector<float> got(1 + 68 * 18 * 9);
        cudaMemcpy(got.data(), o, got.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int p = 0; p < 9; ++p) for (int ry = 0; ry < 18; ++ry) for (int rx = 0; rx < boxw; ++rx) {
            const int gx = L.x0 - 3 + rx, gy = L.y0 - 1 + ry;
            const float want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[(((size_t)L.b * P + p) * H + gy) * W + gx] : 0.f;
            if (got[1 + (p * 18 + ry) * boxw + rx] != want) ++bad;
        }
        printf("case %d: spins %.0f mismatches %d\n", t, got[0], bad);
        bad_total += bad;
    }
    printf(bad_total ? "TMA PROBE FAILED\n" : "TMA PROBE OK\n");
    return bad_total ? 3 : 0;
}
