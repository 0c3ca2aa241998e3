// kintinuous_b200 -- conservative bounding box (in LOGICAL voxel coordinates) of the voxels a depth frame can update.
// CUDA-free so that tests/test_frustum_box.py can check it on the CPU (tests/cpp/frustum_host.cpp).
//
// Not part of the reference: tsdf23 (cuda/tsdf_volume.cu:541-640) launches one thread per (x, y) column of the WHOLE volume and lets
// each voxel fail its own projection test.  A voxel can be updated only if its centre g projects inside the image with positive depth:
//   p = Rinv (g - t),  1 / p_z >= 0,  -0.5 <= fx p_x / p_z + cx < cols - 0.5,  -0.5 <= fy p_y / p_z + cy < rows - 0.5      (:579-589)
// i.e. g lies in the pyramid with apex t spanned by the four image-corner rays.  Inside the cube [0, size]^3 that pyramid is covered by
// the convex hull of t and the four corner-ray points at camera depth Zmax = the largest camera depth of any corner of the cube, so the
// axis-aligned box of those five points (clipped to the cube, widened by a margin) contains every voxel the kernel could touch.  The
// launch grid of integrate_kernel is restricted to that box: for a camera at the centre of a 6 m cube it is ~15 % of the columns, and
// the ~85 % of the CTAs whose threads would all have returned from the per-column frustum test are never scheduled.
// Margin: pixel rectangle widened by 2 px on every side, box widened by 3 voxels on every side (the per-column test inside the kernel
// stays in place; this box only has to be a superset).
#pragma once
#include <cmath>

namespace kt {

struct VoxelBox { int lo[3], hi[3]; bool empty; };      // inclusive logical voxel ranges

// Rinv: row-major inverse camera rotation (volume -> camera), t: camera position in the volume frame (metres), cell: voxel size,
// k4 = fx, fy, cx, cy.
inline VoxelBox frustum_voxel_box(const float* Rinv, const float* t, const float* k4, int rows, int cols, int V, const float* cell)
{
    VoxelBox b;
    b.empty = false;
    // R = Rinv^-1 in double (Rinv is a float rotation inverse: its transpose is the inverse only to ~1e-7; invert properly)
    const double a[9] = {Rinv[0], Rinv[1], Rinv[2], Rinv[3], Rinv[4], Rinv[5], Rinv[6], Rinv[7], Rinv[8]};
    const double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
    bool ok = std::isfinite(det) && std::fabs(det) > 1e-6;
    double R[9];
    if (ok) {
        const double id = 1.0 / det;
        R[0] = (a[4] * a[8] - a[5] * a[7]) * id; R[1] = (a[2] * a[7] - a[1] * a[8]) * id; R[2] = (a[1] * a[5] - a[2] * a[4]) * id;
        R[3] = (a[5] * a[6] - a[3] * a[8]) * id; R[4] = (a[0] * a[8] - a[2] * a[6]) * id; R[5] = (a[2] * a[3] - a[0] * a[5]) * id;
        R[6] = (a[3] * a[7] - a[4] * a[6]) * id; R[7] = (a[1] * a[6] - a[0] * a[7]) * id; R[8] = (a[0] * a[4] - a[1] * a[3]) * id;
    }
    for (int i = 0; i < 3; ++i) ok = ok && std::isfinite((double)t[i]) && cell[i] > 0.f;
    ok = ok && k4[0] > 0.f && k4[1] > 0.f;
    if (!ok) {                                   // degenerate input: no restriction
        for (int i = 0; i < 3; ++i) { b.lo[i] = 0; b.hi[i] = V - 1; }
        return b;
    }
    // largest camera depth of the cube's corners (voxel centres lie strictly inside the cube)
    double zmax = -1e300;
    for (int c = 0; c < 8; ++c) {
        const double g[3] = {(c & 1) ? (double)V * cell[0] : 0.0, (c & 2) ? (double)V * cell[1] : 0.0, (c & 4) ? (double)V * cell[2] : 0.0};
        const double pz = a[6] * (g[0] - t[0]) + a[7] * (g[1] - t[1]) + a[8] * (g[2] - t[2]);
        if (pz > zmax) zmax = pz;
    }
    if (!(zmax > 0.0)) { b.empty = true; for (int i = 0; i < 3; ++i) { b.lo[i] = 0; b.hi[i] = -1; } return b; }   // the whole cube is behind the camera
    zmax *= 1.001;
    double lo[3] = {t[0], t[1], t[2]}, hi[3] = {t[0], t[1], t[2]};
    const double us[2] = {-2.5, (double)cols + 1.5}, vs[2] = {-2.5, (double)rows + 1.5};
    for (int iu = 0; iu < 2; ++iu)
        for (int iv = 0; iv < 2; ++iv) {
            const double pc[3] = {zmax * (us[iu] - k4[2]) / k4[0], zmax * (vs[iv] - k4[3]) / k4[1], zmax};
            for (int i = 0; i < 3; ++i) {
                const double g = t[i] + R[i * 3 + 0] * pc[0] + R[i * 3 + 1] * pc[1] + R[i * 3 + 2] * pc[2];
                if (g < lo[i]) lo[i] = g;
                if (g > hi[i]) hi[i] = g;
            }
        }
    for (int i = 0; i < 3; ++i) {
        // voxel x has its centre at (x + 0.5) * cell
        double l = std::floor(lo[i] / cell[i] - 0.5) - 3.0, h = std::ceil(hi[i] / cell[i] - 0.5) + 3.0;
        if (l < 0.0) l = 0.0;
        if (h > (double)(V - 1)) h = (double)(V - 1);
        if (l > h) { b.empty = true; b.lo[i] = 0; b.hi[i] = -1; }
        else { b.lo[i] = (int)l; b.hi[i] = (int)h; }
    }
    if (b.empty) for (int i = 0; i < 3; ++i) { b.lo[i] = 0; b.hi[i] = -1; }
    return b;
}

// Tiles of `tile` storage coordinates covering the logical range [lo, hi] shifted by `wrap` (0 <= wrap < V) on a cyclic axis of V
// (V % tile == 0): first tile index and tile count (<= V / tile).
inline void cyclic_tile_range(int lo, int hi, int wrap, int V, int tile, int* first_tile, int* n_tiles)
{
    const int tiles = V / tile;
    const int s0 = (lo + wrap) % V;
    const int len = hi - lo + 1;
    const int t0 = s0 / tile;
    int n = (s0 - t0 * tile + len + tile - 1) / tile;
    if (n > tiles) n = tiles;
    *first_tile = t0; *n_tiles = n;
}

} // namespace kt
