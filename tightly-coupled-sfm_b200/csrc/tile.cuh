// 64x16-pixel output tiles with a halo ring, 256 threads, 4 pixels per thread.
//
// Shared-memory tiles are indexed in *padded-image* coordinates: cell (cx, cy),
// cx in [-R, 64+R), cy in [-R, 16+R), holds the value at image position
// (reflect1(y0+cy), reflect1(x0+cx)) -- i.e. ReflectionPad2d(1) is applied when
// the tile is filled, so a 3x3 window is a plain neighbourhood read.  Only one
// pixel of padding exists; cells further outside the image are never read by a
// consumer and are filled with zeros.
#pragma once
#include "tcsfm_math.cuh"

namespace tcsfm {

constexpr int kTileW = 64;
constexpr int kTileH = 16;
constexpr int kTileThreads = 256;
constexpr int kPixPerThread = 4;    // a vertical strip of 4 rows at one column

// thread `tid`'s k-th own pixel inside the tile: warps span 32 consecutive columns
__device__ __forceinline__ void own_pixel(int tid, int k, int& tx, int& ty) {
    tx = tid & (kTileW - 1);
    ty = (tid >> 6) * kPixPerThread + k;
}

template <int R>
struct Tile {
    static constexpr int kPitch = kTileW + 2 * R;
    static constexpr int kRows = kTileH + 2 * R;
    static constexpr int kCells = kPitch * kRows;
    // linear cell index of tile-relative coordinates (may be negative down to -R)
    __device__ __forceinline__ static int cell(int cx, int cy) { return (cy + R) * kPitch + (cx + R); }
    __device__ __forceinline__ static void cell_xy(int cell, int& cx, int& cy) {
        cy = cell / kPitch;
        cx = cell - cy * kPitch - R;
        cy -= R;
    }
    // Maps a cell to the image pixel whose value it holds.  Returns false for
    // cells that no consumer reads (more than one pixel outside the image).
    __device__ __forceinline__ static bool cell_to_reflected(int cell, int x0, int y0, int H, int W, int& ry, int& rx) {
        int cx, cy;
        cell_xy(cell, cx, cy);
        const int gx = x0 + cx, gy = y0 + cy;
        if (gx < -1 || gx > W || gy < -1 || gy > H) return false;
        rx = reflect1(gx, W);
        ry = reflect1(gy, H);
        return true;
    }
};

}  // namespace tcsfm
