// Minimal 2-D TMA probe: box 64 x 16 floats of a [H][W] tensor.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap map, float* out, int cx, int cy, int variant) {
    __shared__ __align__(128) float cs[64 * 16];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(&bar)), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(&bar)), "r"(64u * 16u * 4u) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     :: "r"(smem_addr(cs)), "l"(&map), "r"(smem_addr(&bar)), "r"(cx), "r"(cy) : "memory");
    }
    unsigned spins = 0;
    for (;; ++spins) {
        unsigned done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_addr(&bar)), "r"(0u) : "memory");
        if (done) break;
        if (spins > (1u << 22)) { if (threadIdx.x == 0) out[0] = -12345.f; return; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 16; i += blockDim.x) out[1 + i] = cs[i];
    if (threadIdx.x == 0) out[0] = (float)spins;
}
int main(int argc, char** argv) {
    const int H = 64, W = 96;
    std::vector<float> h((size_t)H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i + 1.0f;
    float *d, *o;
    cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&o, (1 + 64 * 16) * 4); cudaMemset(o, 0, (1 + 64 * 16) * 4);
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    CUtensorMap map; memset(&map, 0, sizeof(map));
    const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
    const cuuint64_t strides[1] = {(cuuint64_t)W * 4};
    const cuuint32_t box[2] = {64, 16}, estr[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)fp)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode2d: %d\n", (int)r);
    int dev = 0; cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev); printf("device %s cc %d.%d\n", pr.name, pr.major, pr.minor);
    const int cx = argc > 1 ? atoi(argv[1]) : 0, cy = argc > 2 ? atoi(argv[2]) : 0;
    probe<<<1, 128>>>(map, o, cx, cy, 0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("sync: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 2;
    std::vector<float> got(1 + 64 * 16);
    cudaMemcpy(got.data(), o, got.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int ry = 0; ry < 16; ++ry) for (int rx = 0; rx < 64; ++rx) {
        const int gx = cx + rx, gy = cy + ry;
        const float want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[(size_t)gy * W + gx] : 0.f;
        if (got[1 + ry * 64 + rx] != want) ++bad;
    }
    printf("spins %.0f mismatches %d\n", got[0], bad);
    return bad ? 3 : 0;
}
