// Fused pair loss: Compute_Loss.compute_pairwise_loss (reference losses.py:151-183)
// + the sums of mean_on_mask (losses.py:142-149), forward and backward.
//
// Forward, per 64x16 tile of one pair: every thread back-projects / projects its
// 4 pixels (and the CTA the 1-pixel halo ring), bilinear-gathers the source image
// into a shared-memory tile next to the coalesced target tile, then evaluates
// L1, the auto-mask, the 3x3 SSIM and the depth-consistency weight from shared
// memory and block-reduces the three masked sums (one atomic each per CTA).
// Nothing but diff_img / mask (4 B/px each) is written.
//
// Backward recomputes the warped tile with a 2-pixel halo instead of saving it:
// SSIM adjoint coefficients are built for the +1 ring channel by channel, each
// own pixel gathers its 3x3 coefficient neighbourhood (reflection handled by tap
// multiplicities), adds the L1 / depth-consistency adjoints and pushes the result
// through the bilinear + projective adjoint: grad(target depth) is a direct
// store, grad(source depth) a 4-tap atomic scatter, grad(K[R|t]) 12 block-reduced
// accumulators per batch element.
#include "tile.cuh"

namespace tcsfm {

constexpr int kMaxGroups = 8;

struct PairLaunch {
    tcsfm_pair_group g[kMaxGroups];
    Arith A;
    float w_l1, w_ssim, C1, C2;
    int flags;
};

struct PairCtx {
    const float* tgt; const float* ref; const float* tdep; const float* rdep;
    int64_t tgt_sc, ref_sc;
};

// Fills one shared-memory cell of the target / warped tiles for an image pixel
// (ry, rx) (already reflected).  Returns the geometry in `p`.
__device__ __forceinline__ void fill_cell(const PairCtx& c, const Cam& cam, const Arith& A, int rx, int ry,
                                          float* ts, float* ws, int cells, int cell, WarpPt& p) {
    const int pix = ry * A.W + rx;
    warp_point(cam, A, rx, ry, __ldg(c.tdep + pix), p);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        ws[ch * cells + cell] = sample_plane(c.ref + ch * c.ref_sc, p, A.H, A.W);
        ts[ch * cells + cell] = __ldg(c.tgt + ch * c.tgt_sc + pix);
    }
}

__device__ __forceinline__ void zero_cell(float* ts, float* ws, int cells, int cell) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) { ws[ch * cells + cell] = 0.f; ts[ch * cells + cell] = 0.f; }
}

__device__ __forceinline__ PairCtx make_ctx(const tcsfm_pair_group& g, int b, int n) {
    PairCtx c;
    c.tgt = g.tgt_img + b * g.tgt_sb;
    c.ref = g.ref_img + b * g.ref_sb;
    c.tdep = g.tgt_depth + (int64_t)b * n;
    c.rdep = g.ref_depth ? g.ref_depth + (int64_t)b * n : nullptr;
    c.tgt_sc = g.tgt_sc;
    c.ref_sc = g.ref_sc;
    return c;
}

__global__ void __launch_bounds__(kTileThreads)
pair_fwd_kernel(const __grid_constant__ PairLaunch L) {
    using T1 = Tile<1>;
    TCSFM_DYN_SMEM(float, smem);
    float* ts = smem;
    float* ws = smem + 3 * T1::kCells;
    TCSFM_SHARED float red[3 * (kTileThreads / 32)];

    const tcsfm_pair_group& g = L.g[blockIdx.z];
    const Arith& A = L.A;
    const int H = A.H, W = A.W, n = H * W;
    const int b = blockIdx.y;
    const int tiles_x = (W + kTileW - 1) / kTileW;
    const int tile_y = blockIdx.x / tiles_x, tile_x = blockIdx.x - tile_y * tiles_x;
    const int x0 = tile_x * kTileW, y0 = tile_y * kTileH;
    const Cam cam = load_cam(g.kinv, g.proj, b);
    const PairCtx c = make_ctx(g, b, n);
    const bool auto_mask = (L.flags & TCSFM_AUTO_MASK) != 0;
    const bool depth_mask = (L.flags & TCSFM_DEPTH_MASK) != 0;
    const bool depth_consist = (L.flags & TCSFM_DEPTH_CONSIST) != 0;
    const bool need_depth = depth_mask || depth_consist;

    float own_mask[kPixPerThread], own_dd[kPixPerThread];
    // ---- phase A: own pixels (geometry results kept in registers) ----
#pragma unroll
    for (int k = 0; k < kPixPerThread; ++k) {
        int tx, ty;
        own_pixel(threadIdx.x, k, tx, ty);
        const int gx = x0 + tx, gy = y0 + ty;
        const int cell = T1::cell(tx, ty);
        float m = 0.f, dd = 0.f;
        if (gx < W && gy < H) {
            WarpPt p;
            fill_cell(c, cam, A, gx, gy, ts, ws, T1::kCells, cell, p);
            m = p.valid ? 1.f : 0.f;
            if (auto_mask) {
                const int pix = gy * W + gx;
                float l1[3], ar[3];
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const float t = ts[ch * T1::kCells + cell];
                    l1[ch] = clamp01_nan(fabsf(__fsub_rn(t, ws[ch * T1::kCells + cell])));
                    ar[ch] = fabsf(__fsub_rn(t, __ldg(c.ref + ch * c.ref_sc + pix)));
                }
                if (!(mean3(l1[0], l1[1], l1[2], A) < mean3(ar[0], ar[1], ar[2], A))) m = 0.f;
            }
            if (need_depth) dd = depth_inconsistency(p.Z, sample_plane(c.rdep, p, H, W));
        } else {
            int ry, rx;
            WarpPt p;
            if (T1::cell_to_reflected(cell, x0, y0, H, W, ry, rx)) fill_cell(c, cam, A, rx, ry, ts, ws, T1::kCells, cell, p);
            else zero_cell(ts, ws, T1::kCells, cell);
        }
        own_mask[k] = m;
        own_dd[k] = dd;
    }
    // ---- phase A': the halo ring ----
    for (int cell = threadIdx.x; cell < T1::kCells; cell += kTileThreads) {
        int cx, cy;
        T1::cell_xy(cell, cx, cy);
        if (cx >= 0 && cx < kTileW && cy >= 0 && cy < kTileH) continue;
        int ry, rx;
        WarpPt p;
        if (T1::cell_to_reflected(cell, x0, y0, H, W, ry, rx)) fill_cell(c, cam, A, rx, ry, ts, ws, T1::kCells, cell, p);
        else zero_cell(ts, ws, T1::kCells, cell);
    }
    __syncthreads();
    // ---- phase C: photometric error per own pixel ----
    float part[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < kPixPerThread; ++k) {
        int tx, ty;
        own_pixel(threadIdx.x, k, tx, ty);
        const int gx = x0 + tx, gy = y0 + ty;
        if (gx < W && gy < H) {
            const int cell = T1::cell(tx, ty);
            float e[3];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const float* tc = ts + ch * T1::kCells + cell;
                const float* wc = ws + ch * T1::kCells + cell;
                const float l1 = clamp01_nan(fabsf(__fsub_rn(*tc, *wc)));
                const float s = ssim_value(tc, wc, T1::kPitch, L.C1, L.C2);
                e[ch] = __fadd_rn(__fmul_rn(l1, L.w_l1), __fmul_rn(s, L.w_ssim));
            }
            float diff = mean3(e[0], e[1], e[2], A);
            if (depth_mask) diff = __fmul_rn(diff, __fsub_rn(1.0f, own_dd[k]));
            const int64_t o = (int64_t)b * n + gy * W + gx;
            if (g.diff_img) g.diff_img[o] = diff;
            if (g.mask) g.mask[o] = own_mask[k];
            part[0] += diff * own_mask[k];
            part[1] += own_mask[k];
            if (depth_consist) part[2] += own_dd[k] * own_mask[k];
        }
    }
    block_atomic_accumulate<3>(part, red, g.sums, threadIdx.x, kTileThreads);
}

// ---------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------
struct BwdScalars { float c_rep, c_dep; };

__device__ __forceinline__ BwdScalars bwd_scalars(const tcsfm_pair_group& g, bool depth_consist) {
    BwdScalars s;
    s.c_rep = 0.f; s.c_dep = 0.f;
    if (g.g_scalars) {
        const float s1 = __ldg(g.sums + 1);
        if (s1 > 10000.0f) {                        // mean_on_mask, losses.py:144
            s.c_rep = __ldg(g.g_scalars + 0) / s1;
            if (depth_consist) s.c_dep = __ldg(g.g_scalars + 1) / s1;
        }
    }
    return s;
}

__global__ void __launch_bounds__(kTileThreads)
pair_bwd_kernel(const __grid_constant__ PairLaunch L) {
    using T2 = Tile<2>;
    using T1 = Tile<1>;
    TCSFM_DYN_SMEM(float, smem);
    float* ts = smem;                          // [3][T2]
    float* ws = ts + 3 * T2::kCells;           // [3][T2]
    float* G1 = ws + 3 * T2::kCells;           // [T1] upstream grad of each channel's ssim value at q
    float* cA = G1 + T1::kCells;               // [T1] coefficients of the current channel
    float* cB = cA + T1::kCells;
    float* cC = cB + T1::kCells;
    float* sv = cC + T1::kCells;               // [T1] ssim value of the current channel (depth-mask only)
    TCSFM_SHARED float red[12 * (kTileThreads / 32)];

    const tcsfm_pair_group& g = L.g[blockIdx.z];
    const Arith& A = L.A;
    const int H = A.H, W = A.W, n = H * W;
    const int b = blockIdx.y;
    const int tiles_x = (W + kTileW - 1) / kTileW;
    const int tile_y = blockIdx.x / tiles_x, tile_x = blockIdx.x - tile_y * tiles_x;
    const int x0 = tile_x * kTileW, y0 = tile_y * kTileH;
    const Cam cam = load_cam(g.kinv, g.proj, b);
    const PairCtx c = make_ctx(g, b, n);
    const bool depth_mask = (L.flags & TCSFM_DEPTH_MASK) != 0;
    const bool depth_consist = (L.flags & TCSFM_DEPTH_CONSIST) != 0;
    const bool need_depth = depth_mask || depth_consist;
    const BwdScalars sc = bwd_scalars(g, depth_consist);
    const float* gdiff = g.g_diff ? g.g_diff + (int64_t)b * n : nullptr;
    const float* mask = g.mask + (int64_t)b * n;

    // ---- phase A: target + re-warped source tiles with a 2-pixel halo, and the
    //      upstream gradient of the per-channel SSIM values on the +1 ring ----
    for (int cell = threadIdx.x; cell < T2::kCells; cell += kTileThreads) {
        int cx, cy, ry, rx;
        T2::cell_xy(cell, cx, cy);
        float g1 = 0.f;
        if (T2::cell_to_reflected(cell, x0, y0, H, W, ry, rx)) {
            WarpPt p;
            fill_cell(c, cam, A, rx, ry, ts, ws, T2::kCells, cell, p);
            const int gx = x0 + cx, gy = y0 + cy;
            if (gx >= 0 && gx < W && gy >= 0 && gy < H) {      // a real output pixel q
                const int pix = gy * W + gx;
                float Gd = sc.c_rep * __ldg(mask + pix);
                if (gdiff) Gd += __ldg(gdiff + pix);
                if (depth_mask) Gd *= (1.0f - depth_inconsistency(p.Z, sample_plane(c.rdep, p, H, W)));
                g1 = Gd * A.third * L.w_ssim;
            }
        } else {
            zero_cell(ts, ws, T2::kCells, cell);
        }
        if (cx >= -1 && cx <= kTileW && cy >= -1 && cy <= kTileH) G1[T1::cell(cx, cy)] = g1;
    }
    __syncthreads();

    float gw[kPixPerThread][3];      // grad wrt the warped image at the own pixels
    float d0[kPixPerThread];         // un-weighted photometric error (depth-mask only)
#pragma unroll
    for (int k = 0; k < kPixPerThread; ++k) { gw[k][0] = gw[k][1] = gw[k][2] = 0.f; d0[k] = 0.f; }

    for (int ch = 0; ch < 3; ++ch) {
        const float* tch = ts + ch * T2::kCells;
        const float* wch = ws + ch * T2::kCells;
        // ---- phase B: SSIM adjoint coefficients of this channel on the +1 ring ----
        for (int cell = threadIdx.x; cell < T1::kCells; cell += kTileThreads) {
            int cx, cy;
            T1::cell_xy(cell, cx, cy);
            const float g1 = G1[cell];
            const int gx = x0 + cx, gy = y0 + cy;
            float a = 0.f, bb = 0.f, cc = 0.f, val = 0.f;
            if (gx >= 0 && gx < W && gy >= 0 && gy < H && (g1 != 0.f || depth_mask)) {
                const int c2 = T2::cell(cx, cy);
                const SsimStats s = ssim_stats(tch + c2, wch + c2, T2::kPitch);
                const SsimTerms t = ssim_terms(s, L.C1, L.C2);
                const SsimCoef k = ssim_coef(s, t, g1);       // x = target, y = warped
                a = k.Ay; bb = k.B; cc = k.Cc;
                val = clamp01_nan(t.raw);
            }
            cA[cell] = a; cB[cell] = bb; cC[cell] = cc; sv[cell] = val;
        }
        __syncthreads();
        // ---- phase C: gather the 3x3 coefficient neighbourhood of each own pixel ----
#pragma unroll
        for (int k = 0; k < kPixPerThread; ++k) {
            int tx, ty;
            own_pixel(threadIdx.x, k, tx, ty);
            const int gx = x0 + tx, gy = y0 + ty;
            if (gx < W && gy < H) {
                float sA = 0.f, sB = 0.f, sC = 0.f;
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy) {
                    const int my = reflect_mult(gy + dy, gy, H);
#pragma unroll
                    for (int dx = -1; dx <= 1; ++dx) {
                        const float m = (float)(my * reflect_mult(gx + dx, gx, W));
                        const int c1 = T1::cell(tx + dx, ty + dy);
                        sA += m * cA[c1]; sB += m * cB[c1]; sC += m * cC[c1];
                    }
                }
                const int c2 = T2::cell(tx, ty);
                const float t = tch[c2], w = wch[c2];
                gw[k][ch] = sA + 2.0f * w * sB + t * sC;
                if (depth_mask) {
                    const float l1 = clamp01_nan(fabsf(__fsub_rn(t, w)));
                    d0[k] += __fadd_rn(__fmul_rn(l1, L.w_l1), __fmul_rn(sv[T1::cell(tx, ty)], L.w_ssim));
                }
            }
        }
        __syncthreads();
    }

    // ---- phase D: L1 / depth adjoints and the geometry adjoint per own pixel ----
    float acc[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) acc[i] = 0.f;
#pragma unroll
    for (int k = 0; k < kPixPerThread; ++k) {
        int tx, ty;
        own_pixel(threadIdx.x, k, tx, ty);
        const int gx = x0 + tx, gy = y0 + ty;
        if (gx < W && gy < H) {
            const int pix = gy * W + gx;
            const int c2 = T2::cell(tx, ty);
            WarpPt p;
            warp_point(cam, A, gx, gy, __ldg(c.tdep + pix), p);
            const float m = __ldg(mask + pix);
            float Gd = sc.c_rep * m;
            if (gdiff) Gd += __ldg(gdiff + pix);
            float pd = 0.f, dd = 0.f;
            Taps td;
            if (need_depth) {
                td = gather_taps(c.rdep, p, H, W);
                pd = bilinear(td, p);
                dd = depth_inconsistency(p.Z, pd);
            }
            float G0 = Gd, Gdd = sc.c_dep * m;
            if (depth_mask) {
                G0 = Gd * (1.0f - dd);
                Gdd -= Gd * (d0[k] * A.third);
            }
            float g_ix = 0.f, g_iy = 0.f;
            const float gl1 = G0 * A.third * L.w_l1;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const float t = ts[ch * T2::kCells + c2], w = ws[ch * T2::kCells + c2];
                const float dlt = t - w;
                float gwc = gw[k][ch];
                if (fabsf(dlt) <= 1.0f) gwc += (dlt > 0.f) ? -gl1 : ((dlt < 0.f) ? gl1 : 0.f);
                const Taps ti = gather_taps(c.ref + ch * c.ref_sc, p, H, W);
                bilinear_grad(ti, p, gwc, g_ix, g_iy);
            }
            float g_Z = 0.f, g_pd = 0.f;
            if (need_depth && Gdd != 0.f) {
                depth_inconsistency_adjoint(p.Z, pd, Gdd, g_Z, g_pd);
                bilinear_grad(td, p, g_pd, g_ix, g_iy);
                if (g.g_ref_depth && g_pd != 0.f) {
                    float* plane = g.g_ref_depth + (int64_t)b * n;
                    const bool x0in = (p.x0 >= 0) && (p.x0 < W), x1in = (p.x0 + 1 >= 0) && (p.x0 + 1 < W);
                    const bool y0in = (p.y0 >= 0) && (p.y0 < H), y1in = (p.y0 + 1 >= 0) && (p.y0 + 1 < H);
                    float* r0 = plane + (int64_t)p.y0 * W + p.x0;
                    if (y0in && x0in) atomicAdd(r0, g_pd * (p.wx0 * p.wy0));
                    if (y0in && x1in) atomicAdd(r0 + 1, g_pd * (p.wx1 * p.wy0));
                    if (y1in && x0in) atomicAdd(r0 + W, g_pd * (p.wx0 * p.wy1));
                    if (y1in && x1in) atomicAdd(r0 + W + 1, g_pd * (p.wx1 * p.wy1));
                }
            }
            const GeomGrad gg = geom_adjoint(cam, A, p, g_ix, g_iy, g_Z);
            if (g.g_tgt_depth) g.g_tgt_depth[(int64_t)b * n + pix] = gg.g_depth;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                acc[i * 4 + 0] += gg.gp[i] * p.cam[0];
                acc[i * 4 + 1] += gg.gp[i] * p.cam[1];
                acc[i * 4 + 2] += gg.gp[i] * p.cam[2];
                acc[i * 4 + 3] += gg.gp[i];
            }
        }
    }
    if (g.g_proj) block_atomic_accumulate<12>(acc, red, g.g_proj + b * 12, threadIdx.x, kTileThreads);
}

static int fill_launch(PairLaunch& L, const tcsfm_pair_group* groups, int n, int B, int H, int W,
                       float w_l1, float w_ssim, int flags, const char* who, bool bwd) {
    if (B <= 0 || H < 2 || W < 2) { set_error("%s: bad shape B=%d H=%d W=%d", who, B, H, W); return 1; }
    if (B > 65535) { set_error("%s: B=%d exceeds 65535", who, B); return 1; }
    if (!(flags & TCSFM_SSIM)) { set_error("%s: the fused pair loss requires TCSFM_SSIM (l_ssim)", who); return 1; }
    const bool need_depth = (flags & (TCSFM_DEPTH_MASK | TCSFM_DEPTH_CONSIST)) != 0;
    for (int i = 0; i < n; ++i) {
        const tcsfm_pair_group& g = groups[i];
        if (!g.tgt_img || !g.ref_img || !g.tgt_depth || !g.kinv || !g.proj || !g.sums) {
            set_error("%s: group %d has a null input pointer", who, i); return 1;
        }
        if (need_depth && !g.ref_depth) { set_error("%s: group %d needs ref_depth for the depth terms", who, i); return 1; }
        if (bwd && !g.mask) { set_error("%s: group %d: backward needs the forward mask", who, i); return 1; }
        L.g[i] = g;
    }
    L.A = make_arith(H, W, flags);
    L.w_l1 = w_l1; L.w_ssim = w_ssim;
    L.C1 = (float)(0.01 * 0.01); L.C2 = (float)(0.03 * 0.03);
    L.flags = flags;
    return 0;
}

}  // namespace tcsfm

using namespace tcsfm;

extern "C" int tcsfm_pair_loss_fwd(const tcsfm_pair_group* groups, int n_groups, int B, int H, int W,
                                   float w_l1, float w_ssim, int flags, void* stream) {
    if (!groups || n_groups <= 0) { set_error("tcsfm_pair_loss_fwd: no groups"); return 1; }
    const size_t smem = 6 * Tile<1>::kCells * sizeof(float);
    const int tiles = ((W + kTileW - 1) / kTileW) * ((H + kTileH - 1) / kTileH);
    for (int base = 0; base < n_groups; base += kMaxGroups) {
        const int n = (n_groups - base < kMaxGroups) ? n_groups - base : kMaxGroups;
        PairLaunch L;
        memset(&L, 0, sizeof(L));
        if (int rc = fill_launch(L, groups + base, n, B, H, W, w_l1, w_ssim, flags, "tcsfm_pair_loss_fwd", false)) return rc;
        for (int i = 0; i < n; ++i) cudaMemsetAsync(L.g[i].sums, 0, 4 * sizeof(float), (cudaStream_t)stream);
        dim3 grid(tiles, B, n), block(kTileThreads);
        TCSFM_LAUNCH(pair_fwd_kernel, grid, block, smem, stream, L);
        if (int rc = check_launch("tcsfm_pair_loss_fwd")) return rc;
    }
    return 0;
}

extern "C" int tcsfm_pair_loss_bwd(const tcsfm_pair_group* groups, int n_groups, int B, int H, int W,
                                   float w_l1, float w_ssim, int flags, void* stream) {
    if (!groups || n_groups <= 0) { set_error("tcsfm_pair_loss_bwd: no groups"); return 1; }
    const size_t smem = (6 * Tile<2>::kCells + 5 * Tile<1>::kCells) * sizeof(float);
#ifndef TCSFM_HOST_EMU
    {   // > 48 KB of dynamic shared memory is opt-in (per device, so set it on every call)
        cudaError_t e = cudaFuncSetAttribute(pair_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("tcsfm_pair_loss_bwd: cannot raise dynamic smem to %zu: %s", smem, cudaGetErrorString(e)); return 2; }
    }
#endif
    const int tiles = ((W + kTileW - 1) / kTileW) * ((H + kTileH - 1) / kTileH);
    for (int base = 0; base < n_groups; base += kMaxGroups) {
        const int n = (n_groups - base < kMaxGroups) ? n_groups - base : kMaxGroups;
        PairLaunch L;
        memset(&L, 0, sizeof(L));
        if (int rc = fill_launch(L, groups + base, n, B, H, W, w_l1, w_ssim, flags, "tcsfm_pair_loss_bwd", true)) return rc;
        for (int i = 0; i < n; ++i) {
            if (L.g[i].g_ref_depth) cudaMemsetAsync(L.g[i].g_ref_depth, 0, (size_t)B * H * W * sizeof(float), (cudaStream_t)stream);
            if (L.g[i].g_proj) cudaMemsetAsync(L.g[i].g_proj, 0, (size_t)B * 12 * sizeof(float), (cudaStream_t)stream);
        }
        dim3 grid(tiles, B, n), block(kTileThreads);
        TCSFM_LAUNCH(pair_bwd_kernel, grid, block, smem, stream, L);
        if (int rc = check_launch("tcsfm_pair_loss_bwd")) return rc;
    }
    return 0;
}
