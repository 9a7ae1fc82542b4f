// The loss reduction of per-frame test-time optimisation (PFT):
// DepthOptimizer.compute_optimization_loss, reference
// optimization_experiments/optimizer.py:45-86, as ONE forward and ONE backward launch over the
// error maps that solve_pose_iteratively(return_errors=True) produced (train_mono.py:84-105):
//
//   fwd half  [S*B,1,H,W]: diff_img, valid_mask, auto_mask_error, weight_mask   (target <- source j)
//   inv half  [S*B,1,H,W]: diff_img, valid_mask, auto_mask,       weight_mask   (source j <- target)
//
//   argmin   (optimizer.py:49-69): per pixel min over the S sources of diff (first index on ties),
//            valid = clamp(sum_j valid_j, 0, 1), auto = [min_j diff_j < min_j auto_err_j],
//            term = sum(diff_min * valid * auto * weight[source 0]) / sum(valid * auto)
//   plain    (optimizer.py:71-73): 0.25 * sum(diff * valid * weight) / sum(valid)
//   inverse  (optimizer.py:75-81): 0.25 * sum(diff * valid * weight [* auto]) / sum(valid [* auto])
//   depth    (optimizer.py:83-86): w * mean(1 - weight) over the fwd maps (+ the inv maps)
//
// The backward routes the upstream scalar to diff_img / weight_mask of both halves (masks and
// auto_mask_error are comparisons: no gradient), torch.min's rule for the arg-min.
#include "common.cuh"

namespace tcsfm {

struct PftArgs {
    const float* f_diff; const float* f_valid; const float* f_aerr; const float* f_weight;
    const float* i_diff; const float* i_valid; const float* i_auto; const float* i_weight;
    int B, S;
    int64_t n;                // H*W
    int flags;
    float w_depth;
    float* sums;              // [8]: fwd numerator, fwd mask, inv numerator, inv mask, sum(1-w_fwd), sum(1-w_inv), block counter, -
    float* loss;              // [1]
    const float* g_loss;      // [1] upstream (backward)
    float* g_f_diff; float* g_f_weight; float* g_i_diff; float* g_i_weight;
};

constexpr int kPftThreads = 256;

// Per pixel of batch element b: the arg-min over the sources and the mask of optimizer.py:49-68.
struct PftMin { float diff; int idx; float vmin; };

// kS > 0: the number of sources is a compile-time constant, so the loops unroll and every load of a pixel is issued
// before the first comparison (a run-time bound serialises them: one DRAM round trip per source)
template <int kS>
__device__ __forceinline__ PftMin pft_min(const PftArgs& P, int64_t at, int64_t src_stride, bool automask) {
    const int S = kS > 0 ? kS : P.S;
    PftMin r;
    r.diff = __ldg(P.f_diff + at);
    r.idx = 0;
    float vsum = __ldg(P.f_valid + at);
    float amin = automask ? __ldg(P.f_aerr + at) : 0.f;
#pragma unroll
    for (int j = 1; j < S; ++j) {
        const int64_t o = at + j * src_stride;
        const float v = __ldg(P.f_diff + o);
        // torch.min(dim): the first index holding the minimum wins; a NaN is the minimum
        if (v < r.diff || (v != v && r.diff == r.diff)) { r.diff = v; r.idx = j; }
        vsum += __ldg(P.f_valid + o);
        if (automask) {
            const float a = __ldg(P.f_aerr + o);
            if (a < amin || (a != a && amin == amin)) amin = a;
        }
    }
    r.vmin = vsum < 0.f ? 0.f : (vsum > 1.f ? 1.f : vsum);         // .clamp(0, 1)
    if (automask) r.vmin = (r.diff < amin) ? r.vmin : 0.f;
    return r;
}

template <int kS>
__global__ void __launch_bounds__(kPftThreads)
pft_reduce_fwd_kernel(const PftArgs P) {
    const int S = kS > 0 ? kS : P.S;
    TCSFM_SHARED float red[6 * (kPftThreads / 32)];
    TCSFM_SHARED int last_block;
    const bool argmin = (P.flags & TCSFM_PFT_ARGMIN) != 0, automask = (P.flags & TCSFM_PFT_AUTOMASK) != 0;
    const bool inverse = (P.flags & TCSFM_PFT_INVERSE) != 0, depth = (P.flags & TCSFM_PFT_DEPTH_CONSIST) != 0;
    const int64_t total = (int64_t)P.B * P.n, src_stride = total;
    float part[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int64_t at = (int64_t)blockIdx.x * kPftThreads + threadIdx.x; at < total; at += (int64_t)gridDim.x * kPftThreads) {
        if (argmin) {
            const PftMin m = pft_min<kS>(P, at, src_stride, automask);
            part[0] += m.diff * m.vmin * __ldg(P.f_weight + at);       // weight_mask[0:B]: the first source's
            part[1] += m.vmin;
        }
#pragma unroll
        for (int j = 0; j < S; ++j) {
            const int64_t o = at + j * src_stride;
            const float wf = __ldg(P.f_weight + o);
            if (!argmin) {
                const float v = __ldg(P.f_valid + o);
                part[0] += __ldg(P.f_diff + o) * v * wf;
                part[1] += v;
            }
            if (depth) part[4] += 1.0f - wf;
            if (inverse) {
                const float wi = __ldg(P.i_weight + o);
                float v = __ldg(P.i_valid + o);
                if (automask) v *= __ldg(P.i_auto + o);
                part[2] += __ldg(P.i_diff + o) * wi * v;
                part[3] += v;
                if (depth) part[5] += 1.0f - wi;
            }
        }
    }
    block_atomic_accumulate<6>(part, red, P.sums, threadIdx.x, kPftThreads);
    // the last block to arrive turns the six sums into the loss (optimizer.py:69,73,79-86)
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned done = atomicAdd(reinterpret_cast<unsigned*>(P.sums + 6), 1u);
        last_block = (done == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (last_block && threadIdx.x == 0) {
        __threadfence();
        volatile const float* s = P.sums;
        const float count = (float)((int64_t)S * total);
        float loss = argmin ? s[0] / s[1] : (0.25f * s[0]) / s[1];
        if (inverse) loss += (0.25f * s[2]) / s[3];
        if (depth) {
            loss += P.w_depth * (s[4] / count);
            if (inverse) loss += P.w_depth * (s[5] / count);
        }
        P.loss[0] = loss;
    }
}

__global__ void __launch_bounds__(kPftThreads)
pft_reduce_bwd_kernel(const PftArgs P) {
    const bool argmin = (P.flags & TCSFM_PFT_ARGMIN) != 0, automask = (P.flags & TCSFM_PFT_AUTOMASK) != 0;
    const bool inverse = (P.flags & TCSFM_PFT_INVERSE) != 0, depth = (P.flags & TCSFM_PFT_DEPTH_CONSIST) != 0;
    const int64_t total = (int64_t)P.B * P.n, src_stride = total;
    const float g = __ldg(P.g_loss);
    const float count = (float)((int64_t)P.S * total);
    const float gf = argmin ? g / __ldg(P.sums + 1) : 0.25f * g / __ldg(P.sums + 1);
    const float gi = inverse ? 0.25f * g / __ldg(P.sums + 3) : 0.f;
    const float gd_f = depth ? -g * P.w_depth / count : 0.f;               // d/d weight of w * mean(1 - weight)
    const float gd_i = (depth && inverse) ? gd_f : 0.f;
    for (int64_t at = (int64_t)blockIdx.x * kPftThreads + threadIdx.x; at < total; at += (int64_t)gridDim.x * kPftThreads) {
        PftMin m;
        m.diff = 0.f; m.idx = -1; m.vmin = 0.f;
        if (argmin) m = pft_min<0>(P, at, src_stride, automask);
        for (int j = 0; j < P.S; ++j) {
            const int64_t o = at + j * src_stride;
            float gdiff, gw;
            if (argmin) {
                gdiff = (j == m.idx) ? gf * m.vmin * __ldg(P.f_weight + at) : 0.f;
                gw = (j == 0) ? gf * m.diff * m.vmin : 0.f;
            } else {
                const float v = __ldg(P.f_valid + o);
                gdiff = gf * v * __ldg(P.f_weight + o);
                gw = gf * v * __ldg(P.f_diff + o);
            }
            P.g_f_diff[o] = gdiff;
            P.g_f_weight[o] = gw + gd_f;
            float gid = 0.f, giw = 0.f;
            if (inverse) {
                float v = __ldg(P.i_valid + o);
                if (automask) v *= __ldg(P.i_auto + o);
                gid = gi * v * __ldg(P.i_weight + o);
                giw = gi * v * __ldg(P.i_diff + o);
            }
            P.g_i_diff[o] = gid;
            P.g_i_weight[o] = giw + gd_i;
        }
    }
}

static int fill_pft(PftArgs& P, const char* who) {
    if (P.B <= 0 || P.S <= 0 || P.n <= 0) { set_error("%s: bad shape B=%d S=%d n=%lld", who, P.B, P.S, (long long)P.n); return 1; }
    const bool automask = (P.flags & TCSFM_PFT_AUTOMASK) != 0, inverse = (P.flags & TCSFM_PFT_INVERSE) != 0;
    const bool argmin = (P.flags & TCSFM_PFT_ARGMIN) != 0;
    if (!P.f_diff || !P.f_valid || !P.f_weight || (argmin && automask && !P.f_aerr)) { set_error("%s: null forward-half map", who); return 1; }
    if (inverse && (!P.i_diff || !P.i_valid || !P.i_weight || (automask && !P.i_auto))) { set_error("%s: null inverse-half map", who); return 1; }
    if (!P.sums) { set_error("%s: null sums buffer", who); return 1; }
    return 0;
}

static int pft_grid(const PftArgs& P) {
    const int64_t blocks = ((int64_t)P.B * P.n + kPftThreads - 1) / kPftThreads;
    const int64_t cap = 148 * 8;                       // one wave of eight resident CTAs per SM, grid-stride beyond
    return (int)(blocks < cap ? blocks : cap);
}

}  // namespace tcsfm

using namespace tcsfm;

extern "C" int tcsfm_pft_reduce_fwd(const float* f_diff, const float* f_valid, const float* f_aerr, const float* f_weight,
                                    const float* i_diff, const float* i_valid, const float* i_auto, const float* i_weight,
                                    int B, int S, int64_t n, int flags, float w_depth, float* sums, float* loss, void* stream) {
    PftArgs P;
    memset(&P, 0, sizeof(P));
    P.f_diff = f_diff; P.f_valid = f_valid; P.f_aerr = f_aerr; P.f_weight = f_weight;
    P.i_diff = i_diff; P.i_valid = i_valid; P.i_auto = i_auto; P.i_weight = i_weight;
    P.B = B; P.S = S; P.n = n; P.flags = flags; P.w_depth = w_depth; P.sums = sums; P.loss = loss;
    if (int rc = fill_pft(P, "tcsfm_pft_reduce_fwd")) return rc;
    if (!loss) { set_error("tcsfm_pft_reduce_fwd: null loss pointer"); return 1; }
    cudaMemsetAsync(sums, 0, 8 * sizeof(float), (cudaStream_t)stream);
    switch (S) {
        case 1: TCSFM_LAUNCH(pft_reduce_fwd_kernel<1>, dim3(pft_grid(P)), dim3(kPftThreads), 0, stream, P); break;
        case 2: TCSFM_LAUNCH(pft_reduce_fwd_kernel<2>, dim3(pft_grid(P)), dim3(kPftThreads), 0, stream, P); break;
        case 3: TCSFM_LAUNCH(pft_reduce_fwd_kernel<3>, dim3(pft_grid(P)), dim3(kPftThreads), 0, stream, P); break;
        default: TCSFM_LAUNCH(pft_reduce_fwd_kernel<0>, dim3(pft_grid(P)), dim3(kPftThreads), 0, stream, P); break;
    }
    return check_launch("tcsfm_pft_reduce_fwd");
}

extern "C" int tcsfm_pft_reduce_bwd(const float* f_diff, const float* f_valid, const float* f_aerr, const float* f_weight,
                                    const float* i_diff, const float* i_valid, const float* i_auto, const float* i_weight,
                                    int B, int S, int64_t n, int flags, float w_depth, const float* sums, const float* g_loss,
                                    float* g_f_diff, float* g_f_weight, float* g_i_diff, float* g_i_weight, void* stream) {
    PftArgs P;
    memset(&P, 0, sizeof(P));
    P.f_diff = f_diff; P.f_valid = f_valid; P.f_aerr = f_aerr; P.f_weight = f_weight;
    P.i_diff = i_diff; P.i_valid = i_valid; P.i_auto = i_auto; P.i_weight = i_weight;
    P.B = B; P.S = S; P.n = n; P.flags = flags; P.w_depth = w_depth; P.sums = const_cast<float*>(sums); P.g_loss = g_loss;
    P.g_f_diff = g_f_diff; P.g_f_weight = g_f_weight; P.g_i_diff = g_i_diff; P.g_i_weight = g_i_weight;
    if (int rc = fill_pft(P, "tcsfm_pft_reduce_bwd")) return rc;
    if (!g_loss || !g_f_diff || !g_f_weight || !g_i_diff || !g_i_weight) { set_error("tcsfm_pft_reduce_bwd: null gradient pointer"); return 1; }
    TCSFM_LAUNCH(pft_reduce_bwd_kernel, dim3(pft_grid(P)), dim3(kPftThreads), 0, stream, P);
    return check_launch("tcsfm_pft_reduce_bwd");
}
