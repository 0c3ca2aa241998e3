// kintinuous_b200 -- per-pixel building blocks of the front end, written once and used twice: by the fused per-frame kernel
// (kt_frontend.cu, sources = shared-memory tiles) and by the operator-level kernels behind kt_op_* (kt_pyramid.cu / kt_rgb.cu, sources =
// global memory).  `Src` is any callable src(y, x) returning the finer level's pixel.  Each function states the reference lines whose
// RESULT it reproduces bit for bit; the loops are this repository's own form (fully unrolled 5x5 / 3x3 windows with clipping predicates,
// interior fast paths with compile-time weights, contractions written out as read off the reference build's SASS).
#pragma once
#include "kt_ops.h"

namespace kt {

// global-memory source for the operator-level kernels
template <class T> struct GlobalSrc {
    const T* __restrict__ base; int pitch;
    __device__ __forceinline__ T operator()(int y, int x) const { return base[(size_t)y * pitch + x]; }
};

// pyrDownGaussKernel (bilateral_pyrdown.cu:102-136): edge-aware 5x5 {.375, .25, .0625} decimation of the filtered depth.
// src(y, x): the finer level; (x, y): destination pixel; all products are exact in float (dyadic weights x 16-bit integers).
template <class Src>
__device__ __forceinline__ uint16_t pyrdown_depth_px(const Src& src, int x, int y, int srows, int scols)
{
    const float sigma_color3 = 3 * 30.f;                   // 3 * sigma_color (bilateral_pyrdown.cu:56,121)
    const int center = src(2 * y, 2 * x);
    const int x_mi = max(0, 2 * x - 2) - 2 * x, y_mi = max(0, 2 * y - 2) - 2 * y;
    const int x_ma = min(scols, 2 * x + 3) - 2 * x, y_ma = min(srows, 2 * y + 3) - 2 * y;
    const float weights[3] = {0.375f, 0.25f, 0.0625f};
    float sum = 0, wall = 0;
    if (x_mi == -2 && y_mi == -2 && x_ma == 3 && y_ma == 3) {
        // interior: the whole 5x5 window, no clipping tests (same taps, same order)
#pragma unroll
        for (int yi = -2; yi <= 2; ++yi)
#pragma unroll
            for (int xi = -2; xi <= 2; ++xi) {
                const int val = src(2 * y + yi, 2 * x + xi);
                if (abs(val - center) < sigma_color3) {
                    const float w = weights[xi < 0 ? -xi : xi] * weights[yi < 0 ? -yi : yi];
                    sum += val * w;
                    wall += w;
                }
            }
        return (uint16_t)static_cast<int>(sum / wall);
    }
#pragma unroll
    for (int yi = -2; yi <= 2; ++yi)
#pragma unroll
        for (int xi = -2; xi <= 2; ++xi) {
            if (yi < y_mi || yi >= y_ma || xi < x_mi || xi >= x_ma) continue;
            const int val = src(2 * y + yi, 2 * x + xi);
            if (abs(val - center) < sigma_color3) {
                const float w = weights[xi < 0 ? -xi : xi] * weights[yi < 0 ? -yi : yi];
                sum += val * w;
                wall += w;
            }
        }
    return (uint16_t)static_cast<int>(sum / wall);
}

// pyrDownKernelGaussF / pyrDownKernelIntensityGauss (bilateral_pyrdown.cu:172-233): {1,4,6,4,1}^2 decimation whose window
// [max(0, 2x-2), min(2x+3, scols-1)) excludes the last column / row (Q1), whose weight index runs from the clipped END of the window,
// and whose weight sum is accumulated in an int (Q2).
__device__ __forceinline__ int gauss5(int r, int c)
{
    const int g[5] = {1, 4, 6, 4, 1};
    return g[r] * g[c];
}
template <class Src>
__device__ __forceinline__ float pyrdown_float_px(const Src& src, int x, int y, int srows, int scols)
{
    const int tx = min(2 * x + 3, scols - 1), ty = min(2 * y + 3, srows - 1);
    const int cx0 = max(0, 2 * x - 2), cy0 = max(0, 2 * y - 2);
    float sum = 0; int count = 0;
    if (tx == 2 * x + 3 && ty == 2 * y + 3 && cx0 == 2 * x - 2 && cy0 == 2 * y - 2) {
        // interior: the whole window, weights known at compile time (index (4 - dy, 4 - dx): the table is symmetric)
        const int g5[5] = {1, 4, 6, 4, 1};
#pragma unroll
        for (int dy = 0; dy < 5; ++dy)
#pragma unroll
            for (int dx = 0; dx < 5; ++dx) {
                const float v = src(cy0 + dy, cx0 + dx);
                if (!isnan(v)) { sum = __fmaf_rn(v, (float)(g5[4 - dy] * g5[4 - dx]), sum); count += g5[4 - dy] * g5[4 - dx]; }
            }
        return (float)(sum / (float)count);
    }
#pragma unroll
    for (int dy = 0; dy < 5; ++dy)
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) {
            const int cy = cy0 + dy, cxx = cx0 + dx;
            if (cy >= ty || cxx >= tx) continue;
            const float v = src(cy, cxx);
            if (!isnan(v)) {
                const int g = gauss5(ty - cy - 1, tx - cxx - 1);
                sum = __fmaf_rn(v, (float)g, sum);
                count += g;
            }
        }
    return (float)(sum / (float)count);
}
template <class Src>
__device__ __forceinline__ uint8_t pyrdown_uchar_px(const Src& src, int x, int y, int srows, int scols)
{
    const int tx = min(2 * x + 3, scols - 1), ty = min(2 * y + 3, srows - 1);
    const int cx0 = max(0, 2 * x - 2), cy0 = max(0, 2 * y - 2);
    float sum = 0; int count = 0;
    if (tx == 2 * x + 3 && ty == 2 * y + 3 && cx0 == 2 * x - 2 && cy0 == 2 * y - 2) {
        // interior: 25 taps with compile-time weights; every partial sum is an integer below 2^24, so integer accumulation gives the
        // same float the reference's float accumulation does
        const int g5[5] = {1, 4, 6, 4, 1};
        int isum = 0;
#pragma unroll
        for (int dy = 0; dy < 5; ++dy)
#pragma unroll
            for (int dx = 0; dx < 5; ++dx) isum += (int)src(cy0 + dy, cx0 + dx) * (g5[4 - dy] * g5[4 - dx]);
        return (uint8_t)((float)isum / 256.f);
    }
#pragma unroll
    for (int dy = 0; dy < 5; ++dy)
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) {
            const int cy = cy0 + dy, cxx = cx0 + dx;
            if (cy >= ty || cxx >= tx) continue;
            const int g = gauss5(ty - cy - 1, tx - cxx - 1);
            sum += (float)((int)src(cy, cxx) * g);                   // <= 255 * 36: exact
            count += g;
        }
    return (uint8_t)(sum / (float)count);
}

// applyKernel (bilateral_pyrdown.cu:274-298): the 3x3 gradient pair; the tap index walks 8..0 over the taps actually visited
__device__ __forceinline__ float gsx_tap(int k) { const float t[9] = {0.52201f, 0.00000f, -0.52201f, 0.79451f, -0.00000f, -0.79451f, 0.52201f, 0.00000f, -0.52201f}; return t[k]; }
__device__ __forceinline__ float gsy_tap(int k) { const float t[9] = {0.52201f, 0.79451f, 0.52201f, 0.00000f, 0.00000f, 0.00000f, -0.52201f, -0.79451f, -0.52201f}; return t[k]; }
template <class Src>
__device__ __forceinline__ void gradient_px(const Src& src, int x, int y, int rows, int cols, int16_t& gx, int16_t& gy)
{
    float dxVal = 0, dyVal = 0;
    const int j0 = max(y - 1, 0), j1 = min(y + 1, rows - 1), i0 = max(x - 1, 0), i1 = min(x + 1, cols - 1);
    if (j0 == y - 1 && j1 == y + 1 && i0 == x - 1 && i1 == x + 1) {
        int k = 8;
#pragma unroll
        for (int j = -1; j <= 1; ++j)
#pragma unroll
            for (int i = -1; i <= 1; ++i) {
                const float v = (float)src(y + j, x + i);
                dxVal = __fmaf_rn(v, gsx_tap(k), dxVal);
                dyVal = __fmaf_rn(v, gsy_tap(k), dyVal);
                --k;
            }
    } else {
        const float tx9[9] = {0.52201f, 0.00000f, -0.52201f, 0.79451f, -0.00000f, -0.79451f, 0.52201f, 0.00000f, -0.52201f};
        const float ty9[9] = {0.52201f, 0.79451f, 0.52201f, 0.00000f, 0.00000f, 0.00000f, -0.52201f, -0.79451f, -0.52201f};
        int k = 8;
        for (int j = j0; j <= j1; ++j)
            for (int i = i0; i <= i1; ++i) {
                const float v = (float)src(j, i);
                dxVal = __fmaf_rn(v, tx9[k], dxVal);
                dyVal = __fmaf_rn(v, ty9[k], dyVal);
                --k;
            }
    }
    gx = (int16_t)dxVal; gy = (int16_t)dyVal;
}

// computeVmapKernel's vertex (maps.cu:57-80) from a depth value.  The reference stores the vertex map and computes normals from it in a
// second kernel, so its products are rounded before the normal's differences are taken; when both happen in registers nvcc fuses
// z * (u - cx) * fx_inv - v00.x  into an FMA (seen on hardware: normals off by an ulp) -- explicit round-to-nearest products and
// differences (vertex_of, diff3) keep the reference's rounding.
__device__ __forceinline__ bool vertex_of(int d, int u, int v, float fx_inv, float fy_inv, float cx, float cy, float3& out)
{
    const float z = d / 1000.f;
    if (z != 0) { out.x = __fmul_rn(__fmul_rn(z, (u - cx)), fx_inv); out.y = __fmul_rn(__fmul_rn(z, (v - cy)), fy_inv); out.z = z; return true; }
    return false;
}
__device__ __forceinline__ float3 diff3(const float3& a, const float3& b) { return make_float3(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)); }

// short2FloatKernel (bilateral_pyrdown.cu:235-245) and bgr2IntensityKernel (:247-259; PixelRGB {r, g, b}), per pixel
__device__ __forceinline__ float depth_to_metres(int raw, int cut_off) { return (raw > cut_off || raw <= 0) ? qnan() : ((float)raw) / 1000.0f; }
__device__ __forceinline__ uint8_t rgb_to_intensity(const uchar3& c)
{
    const int value = __fmaf_rn((float)c.y, 0.587f, __fmaf_rn((float)c.x, 0.114f, __fmul_rn((float)c.z, 0.299f)));
    return (uint8_t)value;
}

} // namespace kt
