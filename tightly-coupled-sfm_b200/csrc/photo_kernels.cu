// The `return_errors` photometric block of solve_pose_iteratively (reference
// train_mono.py:84-92; the same arithmetic is compute_photometric_error,
// optimization_experiments/helpers.py:12-18), fused: for a stack of pairs
//
//   auto_mask_error = mean_c(0.15 * clamp|tgt - src| + 0.85 * SSIM(tgt, src))
//   diff_img        = mean_c(0.15 * clamp|rec - tgt| + 0.85 * SSIM(tgt, rec))
//   auto_mask       = diff_img < auto_mask_error
//   weight_mask     = 1 - clamp(|cd - pd| / (cd + pd), 0, 1)
//
// and the backward w.r.t. the reconstructed image `rec` and the two depths.  Same tiling
// and register-window SSIM as the pair kernels (csrc/pair_kernels.cu), minus the geometry:
// the reconstructed image is an input here because the pose network consumes it too.
#include "tile.cuh"

namespace tcsfm {

constexpr int kPhotoCoefPlanes = 9;

struct PhotoArgs {
    const float* tgt; int64_t tgt_sb, tgt_sc;
    const float* src; int64_t src_sb, src_sc;
    const float* rec;                 // [N,3,H,W] contiguous
    const float* pd; const float* cd; // [N,1,H,W]
    const float* valid;               // [N,1,H,W] or NULL: folded into auto_mask (helpers.py:18)
    float* auto_err; float* diff; float* auto_mask; float* weight;
    float* coef;                      // [N,9,H,W] workspace (fwd writes, bwd reads) or NULL
    const float* g_diff; const float* g_weight;
    float* g_rec; float* g_pd; float* g_cd;
    Arith A;
    float w_l1, w_ssim, C1, C2;
};

// one channel of  w_l1 * clamp|x - y| + w_ssim * SSIM(x, y)  down a strip, x/y packed as float2
template <bool kCoef>
__device__ __forceinline__ void strip_channel(const float2* plane, int tx, int ty0, const PhotoArgs& P, bool first,
                                              float (&esum)[kPixPerThread], float* coef_ch, int gx, int gy0, int n) {
    using T1 = Tile<1>;
    float2 v[3][3];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) v[r][cc] = plane[T1::cell(tx - 1 + cc, ty0 - 1 + r)];
#pragma unroll
    for (int k = 0; k < kPixPerThread; ++k) {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) v[(k + 2) % 3][cc] = plane[T1::cell(tx - 1 + cc, ty0 + 1 + k)];
        float2 wv[9], wsq[9];
        float wab[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            wv[i] = v[(k + i / 3) % 3][i % 3];
            wsq[i] = square2_rn(wv[i]);
            wab[i] = __fmul_rn(wv[i].x, wv[i].y);
        }
        const SsimStats s = ssim_stats_packed(wv, wsq, wab);
        const SsimTerms t = ssim_terms(s, P.C1, P.C2);
        const float l1 = clamp01_nan(fabsf(__fsub_rn(wv[4].x, wv[4].y)));
        const float e = __fadd_rn(__fmul_rn(l1, P.w_l1), __fmul_rn(clamp01_nan(t.raw), P.w_ssim));
        esum[k] = first ? e : __fadd_rn(esum[k], e);
        if (kCoef) {
            const int gy = gy0 + k;
            if (coef_ch && gx < P.A.W && gy < P.A.H) {
                const SsimCoef kf = ssim_coef(s, t, P.A.third * P.w_ssim);     // x = target, y = reconstruction
                float* cp = coef_ch + gy * P.A.W + gx;
                cp[0] = kf.Ay; cp[n] = kf.B; cp[2 * (int64_t)n] = kf.Cc;
            }
        }
    }
}

template <int F>
__global__ void __launch_bounds__(kTileThreads, 3)
photo_fwd_kernel(const __grid_constant__ PhotoArgs P) {
    using T1 = Tile<1>;
    TCSFM_DYN_SMEM(float2, tr);                                  // [3][cells] (target, rec)
    float2* tsrc = tr + 3 * T1::kCells;                          // [3][cells] (target, src)
    const Arith& A = P.A;
    const int H = A.H, W = A.W, n = H * W;
    const int b = blockIdx.y;
    const int tiles_x = (W + kTileW - 1) / kTileW;
    const int tile_y = blockIdx.x / tiles_x, tile_x = blockIdx.x - tile_y * tiles_x;
    const int x0 = tile_x * kTileW, y0 = tile_y * kTileH;
    const float* tgt = P.tgt + b * P.tgt_sb;
    const float* src = P.src + b * P.src_sb;
    const float* rec = P.rec + (int64_t)b * 3 * n;
    // two cells per sweep: their eighteen loads are in flight before the first shared-memory store
    for (int cell0 = threadIdx.x; cell0 < T1::kCells; cell0 += 2 * kTileThreads) {
        int pix[2];
        bool ok[2];
        float t[2][3], r[2][3], sv[2][3];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int cell = cell0 + u * kTileThreads;
            int ry = 0, rx = 0;
            ok[u] = cell < T1::kCells && T1::cell_to_reflected(cell, x0, y0, H, W, ry, rx);
            pix[u] = ok[u] ? ry * W + rx : 0;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                t[u][ch] = ok[u] ? __ldg(tgt + ch * P.tgt_sc + pix[u]) : 0.f;
                r[u][ch] = ok[u] ? __ldg(rec + (int64_t)ch * n + pix[u]) : 0.f;
                sv[u][ch] = ok[u] ? __ldg(src + ch * P.src_sc + pix[u]) : 0.f;
            }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int cell = cell0 + u * kTileThreads;
            if (cell < T1::kCells) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    tr[ch * T1::kCells + cell] = make_float2(t[u][ch], r[u][ch]);
                    tsrc[ch * T1::kCells + cell] = make_float2(t[u][ch], sv[u][ch]);
                }
            }
        }
    }
    __syncthreads();
    const int tx = threadIdx.x & (kTileW - 1);
    const int ty0 = (threadIdx.x >> 6) * kPixPerThread;
    const int gx = x0 + tx, gy0 = y0 + ty0;
    float e_rec[kPixPerThread], e_src[kPixPerThread];
    float* coef_b = P.coef ? P.coef + (int64_t)b * kPhotoCoefPlanes * n : nullptr;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        strip_channel<true>(tr + ch * T1::kCells, tx, ty0, P, ch == 0, e_rec, coef_b ? coef_b + (int64_t)3 * ch * n : nullptr, gx, gy0, n);
        strip_channel<false>(tsrc + ch * T1::kCells, tx, ty0, P, ch == 0, e_src, nullptr, gx, gy0, n);
    }
#pragma unroll
    for (int k = 0; k < kPixPerThread; ++k) {
        const int gy = gy0 + k;
        if (gx < W && gy < H) {
            const int64_t o = (int64_t)b * n + gy * W + gx;
            const float diff = mean3_of_sum<F>(e_rec[k], A);
            const float aerr = mean3_of_sum<F>(e_src[k], A);
            if (P.diff) P.diff[o] = diff;
            if (P.auto_err) P.auto_err[o] = aerr;
            if (P.auto_mask) P.auto_mask[o] = (diff < aerr) ? (P.valid ? __ldg(P.valid + o) : 1.f) : 0.f;
            if (P.weight) P.weight[o] = __fsub_rn(1.0f, depth_inconsistency(__ldg(P.cd + o), __ldg(P.pd + o)));
        }
    }
}

__global__ void __launch_bounds__(kTileThreads, 3)
photo_bwd_kernel(const __grid_constant__ PhotoArgs P) {
    using T1 = Tile<1>;
    TCSFM_DYN_SMEM(float, cs);                     // [9][cells] upstream-scaled coefficients
    const Arith& A = P.A;
    const int H = A.H, W = A.W, n = H * W;
    const int b = blockIdx.y;
    const int tiles_x = (W + kTileW - 1) / kTileW;
    const int tile_y = blockIdx.x / tiles_x, tile_x = blockIdx.x - tile_y * tiles_x;
    const int x0 = tile_x * kTileW, y0 = tile_y * kTileH;
    const float* gdiff = P.g_diff ? P.g_diff + (int64_t)b * n : nullptr;
    const float* coef = P.coef + (int64_t)b * kPhotoCoefPlanes * n;
    // two cells per sweep: the twenty loads of both are in flight before the first product is stored
    for (int cell0 = threadIdx.x; cell0 < T1::kCells; cell0 += 2 * kTileThreads) {
        int pix[2];
        bool inside[2];
        float Gd[2], v[2][9];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int cell = cell0 + u * kTileThreads;
            int cx, cy;
            T1::cell_xy(cell < T1::kCells ? cell : 0, cx, cy);
            const int qx = x0 + cx, qy = y0 + cy;
            inside[u] = cell < T1::kCells && gdiff && qx >= 0 && qx < W && qy >= 0 && qy < H;
            pix[u] = inside[u] ? qy * W + qx : 0;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            Gd[u] = inside[u] ? __ldg(gdiff + pix[u]) : 0.f;
#pragma unroll
            for (int j = 0; j < 9; ++j) v[u][j] = inside[u] ? __ldg(coef + (int64_t)j * n + pix[u]) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int cell = cell0 + u * kTileThreads;
            if (cell < T1::kCells) {
#pragma unroll
                for (int j = 0; j < 9; ++j) cs[j * T1::kCells + cell] = Gd[u] * v[u][j];
            }
        }
    }
    __syncthreads();
    const int tx = threadIdx.x & (kTileW - 1);
    const int ty0 = (threadIdx.x >> 6) * kPixPerThread;
    const int gx = x0 + tx;
    const float* tgt = P.tgt + b * P.tgt_sb;
    const float* rec = P.rec + (int64_t)b * 3 * n;
    // the own pixels' target / reconstruction values and upstream: requested before the window sums
    float tq[kPixPerThread][3], wq[kPixPerThread][3], gq[kPixPerThread];
#pragma unroll
    for (int k = 0; k < kPixPerThread; ++k) {
        const int gy = y0 + ty0 + k;
        const bool own = gx < W && gy < H;
        const int pix = own ? gy * W + gx : 0;
        gq[k] = (own && gdiff) ? __ldg(gdiff + pix) : 0.f;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            tq[k][ch] = own ? __ldg(tgt + ch * P.tgt_sc + pix) : 0.f;
            wq[k][ch] = own ? __ldg(rec + (int64_t)ch * n + pix) : 0.f;
        }
    }
    float h[3][9];
    const bool dup_l = (gx == 1), dup_r = (gx == W - 2);
    auto hsum = [&](int r, float (&out)[9]) {
        const int c1 = T1::cell(tx, ty0 - 1 + r);
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const float* pl = cs + j * T1::kCells + c1;
            const float l = pl[-1], m = pl[0], rr = pl[1];
            float s = (l + m) + rr;
            if (dup_l) s += l;
            if (dup_r) s += rr;
            out[j] = s;
        }
    };
    hsum(0, h[0]);
    hsum(1, h[1]);
#pragma unroll
    for (int k = 0; k < kPixPerThread; ++k) {
        hsum(k + 2, h[(k + 2) % 3]);
        const int gy = y0 + ty0 + k;
        if (gx < W && gy < H) {
            const bool dup_u = (gy == 1), dup_d = (gy == H - 2);
            const int pix = gy * W + gx;
            const int64_t o = (int64_t)b * n + pix;
            const float Gd = gq[k];
            const float gl1 = Gd * A.third * P.w_l1;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float V[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int q = 3 * ch + j;
                    float s = (h[k % 3][q] + h[(k + 1) % 3][q]) + h[(k + 2) % 3][q];
                    if (dup_u) s += h[k % 3][q];
                    if (dup_d) s += h[(k + 2) % 3][q];
                    V[j] = s;
                }
                const float t = tq[k][ch], w = wq[k][ch];
                const float dlt = t - w;
                float gw = V[0] + 2.0f * w * V[1] + t * V[2];
                if (fabsf(dlt) <= 1.0f) gw += (dlt > 0.f) ? -gl1 : ((dlt < 0.f) ? gl1 : 0.f);
                if (P.g_rec) P.g_rec[((int64_t)b * 3 + ch) * n + pix] = gw;
            }
            if (P.g_pd || P.g_cd) {
                float g_Z = 0.f, g_p = 0.f;
                if (P.g_weight) depth_inconsistency_adjoint(__ldg(P.cd + o), __ldg(P.pd + o), -__ldg(P.g_weight + o), g_Z, g_p);
                if (P.g_cd) P.g_cd[o] = g_Z;
                if (P.g_pd) P.g_pd[o] = g_p;
            }
        }
    }
}

static int fill_photo(PhotoArgs& P, int N, int H, int W, float w_l1, float w_ssim, int flags, const char* who) {
    if (N <= 0 || H < 2 || W < 2) { set_error("%s: bad shape N=%d H=%d W=%d", who, N, H, W); return 1; }
    if (N > 65535) { set_error("%s: N=%d exceeds 65535", who, N); return 1; }
    if (!P.tgt || !P.rec || !P.pd || !P.cd) { set_error("%s: null input pointer", who); return 1; }
    P.A = make_arith(H, W, flags);
    P.w_l1 = w_l1; P.w_ssim = w_ssim;
    P.C1 = (float)(0.01 * 0.01); P.C2 = (float)(0.03 * 0.03);
    return 0;
}

}  // namespace tcsfm

using namespace tcsfm;

extern "C" int tcsfm_photo_coef_planes(void) { return kPhotoCoefPlanes; }

extern "C" int tcsfm_photo_fwd(const float* tgt, int64_t tgt_sb, int64_t tgt_sc, const float* src, int64_t src_sb, int64_t src_sc,
                               const float* rec, const float* proj_depth, const float* comp_depth,
                               float* auto_err, float* diff, float* auto_mask, float* weight, float* coef,
                               const float* valid,
                               int N, int H, int W, float w_l1, float w_ssim, int flags, void* stream) {
    PhotoArgs P;
    memset(&P, 0, sizeof(P));
    P.tgt = tgt; P.tgt_sb = tgt_sb; P.tgt_sc = tgt_sc; P.src = src; P.src_sb = src_sb; P.src_sc = src_sc;
    P.rec = rec; P.pd = proj_depth; P.cd = comp_depth;
    P.auto_err = auto_err; P.diff = diff; P.auto_mask = auto_mask; P.weight = weight; P.coef = coef; P.valid = valid;
    if (int rc = fill_photo(P, N, H, W, w_l1, w_ssim, flags, "tcsfm_photo_fwd")) return rc;
    if (!src) { set_error("tcsfm_photo_fwd: null src"); return 1; }
    const size_t smem = 6 * Tile<1>::kCells * sizeof(float2);
#ifndef TCSFM_HOST_EMU
    cudaError_t e = cudaSuccess;
    TCSFM_DISPATCH_FLAVOUR(flags, e = cudaFuncSetAttribute(photo_fwd_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (e != cudaSuccess) { set_error("tcsfm_photo_fwd: cannot raise dynamic smem: %s", cudaGetErrorString(e)); return 2; }
#endif
    dim3 grid(((W + kTileW - 1) / kTileW) * ((H + kTileH - 1) / kTileH), N), block(kTileThreads);
    TCSFM_DISPATCH_FLAVOUR(flags, TCSFM_LAUNCH(photo_fwd_kernel<F>, grid, block, smem, stream, P));
    return check_launch("tcsfm_photo_fwd");
}

extern "C" int tcsfm_photo_bwd(const float* tgt, int64_t tgt_sb, int64_t tgt_sc, const float* rec,
                               const float* proj_depth, const float* comp_depth, const float* coef,
                               const float* g_diff, const float* g_weight, float* g_rec, float* g_pd, float* g_cd,
                               int N, int H, int W, float w_l1, float w_ssim, int flags, void* stream) {
    PhotoArgs P;
    memset(&P, 0, sizeof(P));
    P.tgt = tgt; P.tgt_sb = tgt_sb; P.tgt_sc = tgt_sc; P.rec = rec; P.pd = proj_depth; P.cd = comp_depth;
    P.coef = const_cast<float*>(coef); P.g_diff = g_diff; P.g_weight = g_weight; P.g_rec = g_rec; P.g_pd = g_pd; P.g_cd = g_cd;
    if (int rc = fill_photo(P, N, H, W, w_l1, w_ssim, flags, "tcsfm_photo_bwd")) return rc;
    if (!coef) { set_error("tcsfm_photo_bwd: the forward's coef workspace is required"); return 1; }
    const size_t smem = 9 * Tile<1>::kCells * sizeof(float);
    dim3 grid(((W + kTileW - 1) / kTileW) * ((H + kTileH - 1) / kTileH), N), block(kTileThreads);
    TCSFM_LAUNCH(photo_bwd_kernel, grid, block, smem, stream, P);
    return check_launch("tcsfm_photo_bwd");
}
