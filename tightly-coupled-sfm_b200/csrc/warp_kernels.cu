// inverse_warp2 (reference models/stn.py:234-273) forward and backward.
//
// One thread per target pixel: back-project with K^-1 and the depth, apply
// K[R|t], normalise, fill out-of-range coordinates with 2, bilinear-sample the
// source image (3 ch) and source depth (1 ch).  The only HBM traffic is one
// read of depth, the gathers (L1/L2 resident: neighbouring pixels hit the same
// lines) and one write per output; no intermediate tensor is materialised.
#include "tcsfm_math.cuh"

namespace tcsfm {

constexpr int kWarpThreads = 256;
#ifndef TCSFM_WARP_BWD_PIX
#define TCSFM_WARP_BWD_PIX 4
#endif
constexpr int kWarpBwdPix = TCSFM_WARP_BWD_PIX;
#ifndef TCSFM_WARP_BWD_BLOCKS
#define TCSFM_WARP_BWD_BLOCKS 3     // staged loads (every upstream value, then every gather of a pixel in flight together) want 80
#endif                              // registers: 3 CTAs/SM + two pixels unrolled beat the unstaged 5 CTAs/SM at 48 registers by 6-12 %
#ifndef TCSFM_WARP_BWD_UNROLL
#define TCSFM_WARP_BWD_UNROLL 2
#endif
constexpr int kWarpBwdUnroll = TCSFM_WARP_BWD_UNROLL;

// Per-batch-element base pointers are pinned (common.cuh) and everything inside one element is
// addressed with 32-bit offsets (checked by the launchers): one IMAD.WIDE per access.
template <int F>
__global__ void __launch_bounds__(kWarpThreads)
warp_fwd_kernel(const float* __restrict__ img, int64_t img_sb, int img_sc,
                const float* __restrict__ depth, const float* __restrict__ ref_depth,
                const float* __restrict__ kinv, const float* __restrict__ proj,
                float* __restrict__ out_img, float* __restrict__ out_valid,
                float* __restrict__ out_pd, float* __restrict__ out_cd,
                const float* __restrict__ tgt, int64_t tgt_sb, int tgt_sc, float* __restrict__ out_stack, Arith A) {
    const int b = blockIdx.y;
    const int n = A.H * A.W;
    const int pix = blockIdx.x * kWarpThreads + threadIdx.x;
    if (pix >= n) return;
    const Cam c = load_cam(kinv, proj, b);
    const int v = pix / A.W, u = pix - v * A.W;
    const int64_t o = (int64_t)b * n + pix;
    const float* img_b = pin_pointer(img + b * img_sb);
    WarpPt p;
    warp_point<F>(c, A, u, v, __ldg(depth + o), p);
    const TapIdx ti = make_taps(p, A.H, A.W);
    if (out_img || out_stack) {
        float* oimg_b = out_img ? out_img + (int64_t)b * 3 * n : nullptr;
        float* ostk_b = out_stack ? out_stack + (int64_t)b * 6 * n : nullptr;
        const float* tgt_b = out_stack ? tgt + b * tgt_sb : nullptr;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float w = blend(load_taps(img_b, ch * img_sc, ti, A.W), ti);
            if (out_img) oimg_b[ch * n + pix] = w;
            if (out_stack) {          // next pose-net input: [target * valid | reconstruction], train_mono.py:74-76
                ostk_b[(3 + ch) * n + pix] = w;
                ostk_b[ch * n + pix] = __fmul_rn(__ldg(tgt_b + (ch * tgt_sc + pix)), p.valid ? 1.f : 0.f);
            }
        }
    }
    if (out_valid) out_valid[o] = p.valid ? 1.f : 0.f;
    if (out_pd) out_pd[o] = blend(load_taps(ref_depth + (int64_t)b * n, 0, ti, A.W), ti);
    if (out_cd) out_cd[o] = p.Z;
}

template <int F>
__global__ void __launch_bounds__(kWarpThreads, TCSFM_WARP_BWD_BLOCKS)
warp_bwd_kernel(const float* __restrict__ img, int64_t img_sb, int img_sc,
                const float* __restrict__ depth, const float* __restrict__ ref_depth,
                const float* __restrict__ kinv, const float* __restrict__ proj,
                const float* __restrict__ g_oimg, const float* __restrict__ g_opd, const float* __restrict__ g_ocd,
                const float* __restrict__ g_ostack,
                float* __restrict__ g_depth, float* __restrict__ g_ref_depth, float* __restrict__ g_proj,
                float* __restrict__ g_img, Arith A) {
    TCSFM_SHARED float red[12 * (kWarpThreads / 32)];
    const int b = blockIdx.y;
    const int n = A.H * A.W;
    float acc[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) acc[i] = 0.f;
    const Cam c = load_cam(kinv, proj, b);
    const float* img_b = pin_pointer(img + b * img_sb);
    const float* dep_b = pin_pointer(depth + (int64_t)b * n);
    const float* rdep_b = pin_pointer(ref_depth + (int64_t)b * n);
    const float* goimg_b = pin_pointer(g_oimg ? g_oimg + (int64_t)b * 3 * n : nullptr);
    const float* gostk_b = pin_pointer(g_ostack ? g_ostack + (int64_t)b * 6 * n + (int64_t)3 * n : nullptr);
    // kWarpBwdPix pixels per thread: the 12-value block reduction and the camera loads are amortised.  The depth heads
    // the dependent chain depth -> projection -> tap addresses -> gathers: the next pixel's is requested one iteration ahead.
    const int pix0 = blockIdx.x * kWarpBwdPix * kWarpThreads + threadIdx.x;
    float dep_next = pix0 < n ? __ldg(dep_b + pix0) : 1.0f;
#pragma unroll kWarpBwdUnroll
    for (int k = 0; k < kWarpBwdPix; ++k) {
        const int pix = pix0 + k * kWarpThreads;
        if (pix >= n) continue;              // (not `break`: the warp stays provably converged for the reduction's shuffles)
#ifdef TCSFM_WARP_BWD_NO_PREFETCH      // (tuning builds)
        const float dep = __ldg(dep_b + pix);
#else
        const float dep = dep_next;
        if (k + 1 < kWarpBwdPix) dep_next = (pix + kWarpThreads < n) ? __ldg(dep_b + pix + kWarpThreads) : 1.0f;
#endif
        const int v = pix / A.W, u = pix - v * A.W;
#ifndef TCSFM_WARP_BWD_UNSTAGED       // (tuning builds: the per-channel interleaved form)
        // stage 0: every upstream value (none depends on the geometry) is requested before the projection ...
        const bool img_up = goimg_b || gostk_b;
        float gch[3] = {0.f, 0.f, 0.f}, gst[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            if (goimg_b) gch[ch] = __ldg(goimg_b + (ch * n + pix));
            if (gostk_b) gst[ch] = __ldg(gostk_b + (ch * n + pix));
        }
        const float gpd = g_opd ? __ldg(g_opd + (int64_t)b * n + pix) : 0.f;
        const float g_Z = g_ocd ? __ldg(g_ocd + (int64_t)b * n + pix) : 0.f;
        WarpPt p;
        warp_point<F>(c, A, u, v, dep, p);
        const TapIdx ti = make_taps(p, A.H, A.W);
        // ... stage 1: all gathers of the pixel in flight together ...
        Taps tc[3], td;
        if (img_up) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) tc[ch] = load_taps(img_b, ch * img_sc, ti, A.W);
        }
        if (g_opd) td = load_taps(rdep_b, 0, ti, A.W);
        // ... stage 2: the arithmetic and the scatters
        float g_ix = 0.f, g_iy = 0.f;
        if (img_up) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const float g = gch[ch] + gst[ch];
                bilinear_grad(tc[ch], p, g, g_ix, g_iy);
                if (g_img) scatter_taps(g_img + ((int64_t)b * 3 + ch) * n, 0, ti, g, A.W);
            }
        }
        if (g_opd) {
            bilinear_grad(td, p, gpd, g_ix, g_iy);
            if (g_ref_depth) scatter_taps(g_ref_depth + (int64_t)b * n, 0, ti, gpd, A.W);
        }
#else
        WarpPt p;
        warp_point<F>(c, A, u, v, dep, p);
        const TapIdx ti = make_taps(p, A.H, A.W);
        float g_ix = 0.f, g_iy = 0.f;
        if (goimg_b || gostk_b) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float g = goimg_b ? __ldg(goimg_b + (ch * n + pix)) : 0.f;
                if (gostk_b) g += __ldg(gostk_b + (ch * n + pix));
                const Taps t = load_taps(img_b, ch * img_sc, ti, A.W);
                bilinear_grad(t, p, g, g_ix, g_iy);
                if (g_img) scatter_taps(g_img + ((int64_t)b * 3 + ch) * n, 0, ti, g, A.W);
            }
        }
        if (g_opd) {
            const float g = __ldg(g_opd + (int64_t)b * n + pix);
            const Taps t = load_taps(rdep_b, 0, ti, A.W);
            bilinear_grad(t, p, g, g_ix, g_iy);
            if (g_ref_depth) scatter_taps(g_ref_depth + (int64_t)b * n, 0, ti, g, A.W);
        }
        const float g_Z = g_ocd ? __ldg(g_ocd + (int64_t)b * n + pix) : 0.f;
#endif
        const GeomGrad gg = geom_adjoint(c, A, p, g_ix, g_iy, g_Z);
        if (g_depth) g_depth[(int64_t)b * n + pix] = gg.g_depth;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            acc[i * 4 + 0] += gg.gp[i] * p.cam[0];
            acc[i * 4 + 1] += gg.gp[i] * p.cam[1];
            acc[i * 4 + 2] += gg.gp[i] * p.cam[2];
            acc[i * 4 + 3] += gg.gp[i];
        }
    }
    if (g_proj) block_atomic_accumulate<12>(acc, red, g_proj + b * 12, threadIdx.x, kWarpThreads);
}

// in-element offsets are 32-bit in the kernels
static bool strides_fit(int64_t sc, int H, int W) {
    return sc >= 0 && 2 * sc + (int64_t)H * W < ((int64_t)1 << 31) && (int64_t)6 * H * W < ((int64_t)1 << 31);
}

}  // namespace tcsfm

using namespace tcsfm;

extern "C" int tcsfm_warp_fwd(const float* img, int64_t img_sb, int64_t img_sc,
                              const float* depth, const float* ref_depth,
                              const float* kinv, const float* proj,
                              float* out_img, float* out_valid, float* out_proj_depth, float* out_comp_depth,
                              const float* tgt, int64_t tgt_sb, int64_t tgt_sc, float* out_stack,
                              int B, int H, int W, int flags, void* stream) {
    if (B <= 0 || H < 2 || W < 2) { set_error("tcsfm_warp_fwd: bad shape B=%d H=%d W=%d", B, H, W); return 1; }
    if (!img || !depth || !kinv || !proj || (out_proj_depth && !ref_depth)) {
        set_error("tcsfm_warp_fwd: null input pointer"); return 1;
    }
    if (B > 65535) { set_error("tcsfm_warp_fwd: B=%d exceeds 65535", B); return 1; }
    if (out_stack && !tgt) { set_error("tcsfm_warp_fwd: out_stack needs the target image"); return 1; }
    if (!strides_fit(img_sc, H, W) || (tgt && !strides_fit(tgt_sc, H, W))) { set_error("tcsfm_warp_fwd: channel stride / image size out of range"); return 1; }
    const Arith A = make_arith(H, W, flags);
    dim3 grid((H * W + kWarpThreads - 1) / kWarpThreads, B), block(kWarpThreads);
    TCSFM_DISPATCH_FLAVOUR(flags, TCSFM_LAUNCH(warp_fwd_kernel<F>, grid, block, 0, stream, img, img_sb, (int)img_sc, depth,
                                               ref_depth, kinv, proj, out_img, out_valid, out_proj_depth, out_comp_depth,
                                               tgt, tgt_sb, (int)tgt_sc, out_stack, A));
    return check_launch("tcsfm_warp_fwd");
}

extern "C" int tcsfm_warp_bwd(const float* img, int64_t img_sb, int64_t img_sc,
                              const float* depth, const float* ref_depth,
                              const float* kinv, const float* proj,
                              const float* g_out_img, const float* g_out_proj_depth, const float* g_out_comp_depth,
                              const float* g_out_stack,
                              float* g_depth, float* g_ref_depth, float* g_proj, float* g_img,
                              int B, int H, int W, int flags, void* stream) {
    if (B <= 0 || H < 2 || W < 2) { set_error("tcsfm_warp_bwd: bad shape B=%d H=%d W=%d", B, H, W); return 1; }
    if (!img || !depth || !ref_depth || !kinv || !proj) { set_error("tcsfm_warp_bwd: null input pointer"); return 1; }
    if (B > 65535) { set_error("tcsfm_warp_bwd: B=%d exceeds 65535", B); return 1; }
    if (!strides_fit(img_sc, H, W)) { set_error("tcsfm_warp_bwd: channel stride / image size out of range"); return 1; }
    const Arith A = make_arith(H, W, flags);
    const size_t plane = (size_t)B * H * W * sizeof(float);
    if (g_ref_depth) cudaMemsetAsync(g_ref_depth, 0, plane, (cudaStream_t)stream);
    if (g_proj) cudaMemsetAsync(g_proj, 0, (size_t)B * 12 * sizeof(float), (cudaStream_t)stream);
    if (g_img) cudaMemsetAsync(g_img, 0, 3 * plane, (cudaStream_t)stream);
    const int per_block = kWarpThreads * kWarpBwdPix;
    dim3 grid((H * W + per_block - 1) / per_block, B), block(kWarpThreads);
    TCSFM_DISPATCH_FLAVOUR(flags, TCSFM_LAUNCH(warp_bwd_kernel<F>, grid, block, 0, stream, img, img_sb, (int)img_sc, depth,
                                               ref_depth, kinv, proj, g_out_img, g_out_proj_depth, g_out_comp_depth, g_out_stack,
                                               g_depth, g_ref_depth, g_proj, g_img, A));
    return check_launch("tcsfm_warp_bwd");
}
