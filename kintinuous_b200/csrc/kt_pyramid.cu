// kintinuous_b200 -- depth pyramid and vertex / normal maps (sm_100a).
//
// Replaces (reference file:line, src/frontend/cuda/):
//   bilateralFilter / bilateralKernel      bilateral_pyrdown.cu:60-99, :333-343
//   pyrDown / pyrDownGaussKernel           bilateral_pyrdown.cu:102-136, :345-354
//   createVMap / computeVmapKernel         maps.cu:57-80, :123-138
//   createNMap / computeNmapKernel         maps.cu:83-121, :140-155
//   tranformMaps / tranformMapsKernel      maps.cu:157-222
//   resizeVMap/NMap / resizeMapKernel      maps.cu:225-308
// B200 design: the bilateral filter stages a (16+12)x(32+12) depth tile in shared memory once per
// CTA (169 taps/pixel come from smem, not L1); vertex and normal maps of ALL pyramid levels are
// produced by ONE launch straight from the depth pyramid (the vertex map is never re-read to make
// normals); per-pixel arithmetic keeps the reference's expression order (see kt_common.cuh).
// Roofline: HBM-bound streaming except the bilateral filter, which is MUFU(ex2)-bound
// (169 __expf per pixel); algorithmic bytes: DESIGN.md section 4.
#include "kt_ops.h"

namespace kt {

namespace {

const float SIGMA_COLOR = 30.f;      // mm   (bilateral_pyrdown.cu:56)
const float SIGMA_SPACE = 4.5f;      // px   (bilateral_pyrdown.cu:57)

enum { BIL_TX = 32, BIL_TY = 16, BIL_R = 6, BIL_W = BIL_TX + 2 * BIL_R, BIL_H = BIL_TY + 2 * BIL_R };

__global__ void __launch_bounds__(BIL_TX * BIL_TY)
bilateral_kernel(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, int rows, int cols,
                 float sigma_space2_inv_half, float sigma_color2_inv_half)
{
    __shared__ int tile[BIL_H][BIL_W + 1];
    const int x0 = blockIdx.x * BIL_TX, y0 = blockIdx.y * BIL_TY;
    for (int i = threadIdx.y * BIL_TX + threadIdx.x; i < BIL_H * BIL_W; i += BIL_TX * BIL_TY) {
        int ty = i / BIL_W, tx = i - ty * BIL_W;
        int gx = x0 + tx - BIL_R, gy = y0 + ty - BIL_R;
        tile[ty][tx] = (gx >= 0 && gx < cols && gy >= 0 && gy < rows) ? (int)src[(size_t)gy * cols + gx] : 0;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= cols || y >= rows) return;

    const int D = BIL_R * 2 + 1;
    const int value = tile[threadIdx.y + BIL_R][threadIdx.x + BIL_R];
    // interior CTAs (every pixel has its full 13x13 window): same taps in the same order, bounds and the spatial term known at
    // compile time (about 40 % fewer instructions per tap)
    if (x0 >= BIL_R && y0 >= BIL_R && x0 + BIL_TX - 1 + BIL_R + 1 <= cols - 1 && y0 + BIL_TY - 1 + BIL_R + 1 <= rows - 1) {
        float sum1 = 0, sum2 = 0;
#pragma unroll
        for (int dy = -BIL_R; dy <= BIL_R; ++dy) {
            const int* trow = tile[threadIdx.y + BIL_R + dy];
#pragma unroll
            for (int dx = -BIL_R; dx <= BIL_R; ++dx) {
                int tmp = trow[threadIdx.x + BIL_R + dx];
                const float space2 = (float)(dx * dx + dy * dy);
                const float color2 = (float)((value - tmp) * (value - tmp));
                // same contraction as the general loop below compiles to (checked in SASS): fma(space2, k_s, color2 * k_c)
                const float e = __fmaf_rn(space2, sigma_space2_inv_half, __fmul_rn(color2, sigma_color2_inv_half));
                float weight = __expf(-e);
                sum1 = __fmaf_rn((float)tmp, weight, sum1);
                sum2 = __fadd_rn(sum2, weight);
            }
        }
        int res = __float2int_rn(sum1 / sum2);
        dst[(size_t)y * cols + x] = (uint16_t)max(0, min(res, 32767));
        return;
    }
    const int tx = min(x - D / 2 + D, cols - 1);      // exclusive, and clipped to cols-1: Q1
    const int ty = min(y - D / 2 + D, rows - 1);
    float sum1 = 0, sum2 = 0;
    for (int cy = max(y - D / 2, 0); cy < ty; ++cy) {
        const int* trow = tile[cy - y0 + BIL_R];
        for (int cx = max(x - D / 2, 0); cx < tx; ++cx) {
            int tmp = trow[cx - x0 + BIL_R];
            float space2 = (x - cx) * (x - cx) + (y - cy) * (y - cy);
            float color2 = (value - tmp) * (value - tmp);
            float weight = __expf(-(space2 * sigma_space2_inv_half + color2 * sigma_color2_inv_half));
            sum1 += tmp * weight;
            sum2 += weight;
        }
    }
    int res = __float2int_rn(sum1 / sum2);
    dst[(size_t)y * cols + x] = (uint16_t)max(0, min(res, 32767));
}

__global__ void __launch_bounds__(256)
pyrdown_gauss_kernel(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, int srows, int scols, int drows, int dcols, float sigma_color)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dcols || y >= drows) return;
    const int D = 5;
    int center = src[(size_t)(2 * y) * scols + 2 * x];
    int x_mi = max(0, 2 * x - D / 2) - 2 * x;
    int y_mi = max(0, 2 * y - D / 2) - 2 * y;
    int x_ma = min(scols, 2 * x - D / 2 + D) - 2 * x;
    int y_ma = min(srows, 2 * y - D / 2 + D) - 2 * y;
    float sum = 0, wall = 0;
    const float weights[3] = {0.375f, 0.25f, 0.0625f};
    if (x_mi == -2 && y_mi == -2 && x_ma == 3 && y_ma == 3) {
        // interior: the 25 loads are issued together; same taps in the same order (all products of these dyadic weights with
        // a 16-bit integer are exact in float, so the accumulation is bit-identical to the general loop)
        int vals[25];
#pragma unroll
        for (int yi = -2; yi <= 2; ++yi)
#pragma unroll
            for (int xi = -2; xi <= 2; ++xi) vals[(yi + 2) * 5 + xi + 2] = src[(size_t)(2 * y + yi) * scols + 2 * x + xi];
#pragma unroll
        for (int yi = -2; yi <= 2; ++yi)
#pragma unroll
            for (int xi = -2; xi <= 2; ++xi) {
                const int val = vals[(yi + 2) * 5 + xi + 2];
                if (abs(val - center) < 3 * sigma_color) {
                    sum += val * weights[xi < 0 ? -xi : xi] * weights[yi < 0 ? -yi : yi];
                    wall += weights[xi < 0 ? -xi : xi] * weights[yi < 0 ? -yi : yi];
                }
            }
        dst[(size_t)y * dcols + x] = (uint16_t)static_cast<int>(sum / wall);
        return;
    }
    for (int yi = y_mi; yi < y_ma; ++yi)
        for (int xi = x_mi; xi < x_ma; ++xi) {
            int val = src[(size_t)(2 * y + yi) * scols + 2 * x + xi];
            if (abs(val - center) < 3 * sigma_color) {
                sum += val * weights[abs(xi)] * weights[abs(yi)];
                wall += weights[abs(xi)] * weights[abs(yi)];
            }
        }
    dst[(size_t)y * dcols + x] = (uint16_t)static_cast<int>(sum / wall);
}

__device__ __forceinline__ bool vertex_from_depth(const uint16_t* __restrict__ depth, int cols, int u, int v,
                                                  float fx_inv, float fy_inv, float cx, float cy, float3& out)
{
    float z = depth[(size_t)v * cols + u] / 1000.f;        // mm -> m
    if (z != 0) {
        out.x = z * (u - cx) * fx_inv;
        out.y = z * (v - cy) * fy_inv;
        out.z = z;
        return true;
    }
    return false;
}

__global__ void __launch_bounds__(256)
vmap_kernel(const uint16_t* __restrict__ depth, float* __restrict__ vmap, int rows, int cols, float fx_inv, float fy_inv, float cx, float cy)
{
    int u = threadIdx.x + blockIdx.x * blockDim.x;
    int v = threadIdx.y + blockIdx.y * blockDim.y;
    if (u >= cols || v >= rows) return;
    float3 p;
    size_t P = (size_t)rows * cols, i = (size_t)v * cols + u;
    if (vertex_from_depth(depth, cols, u, v, fx_inv, fy_inv, cx, cy, p)) { vmap[i] = p.x; vmap[i + P] = p.y; vmap[i + 2 * P] = p.z; }
    else vmap[i] = qnan();                                 // Q7: only the x plane is invalidated
}

__global__ void __launch_bounds__(256)
nmap_kernel(const float* __restrict__ vmap, float* __restrict__ nmap, int rows, int cols)
{
    int u = threadIdx.x + blockIdx.x * blockDim.x;
    int v = threadIdx.y + blockIdx.y * blockDim.y;
    if (u >= cols || v >= rows) return;
    size_t P = (size_t)rows * cols, i = (size_t)v * cols + u;
    if (u == cols - 1 || v == rows - 1) { nmap[i] = qnan(); return; }
    float3 v00, v01, v10;
    v00.x = vmap[i]; v01.x = vmap[i + 1]; v10.x = vmap[i + cols];
    if (!isnan(v00.x) && !isnan(v01.x) && !isnan(v10.x)) {
        v00.y = vmap[i + P]; v01.y = vmap[i + 1 + P]; v10.y = vmap[i + cols + P];
        v00.z = vmap[i + 2 * P]; v01.z = vmap[i + 1 + 2 * P]; v10.z = vmap[i + cols + 2 * P];
        float3 r = normalized3(cross3(sub3(v01, v00), sub3(v10, v00)));
        nmap[i] = r.x; nmap[i + P] = r.y; nmap[i + 2 * P] = r.z;
    } else nmap[i] = qnan();
}

struct MapsParams { MapsLevel lv[LEVELS]; };

// All levels in one launch: blockIdx.z = level, grid sized for level 0.
__global__ void __launch_bounds__(256)
maps_pyramid_kernel(const MapsParams p)
{
    const MapsLevel& L = p.lv[blockIdx.z];
    const int rows = L.rows, cols = L.cols;
    int u = threadIdx.x + blockIdx.x * blockDim.x;
    int v = threadIdx.y + blockIdx.y * blockDim.y;
    if (u >= cols || v >= rows) return;
    const float fx_inv = L.fx_inv, fy_inv = L.fy_inv, cx = L.k.cx, cy = L.k.cy;   // 1/fx is computed on the HOST (maps.cu:135)
    const size_t P = (size_t)rows * cols, i = (size_t)v * cols + u;
    const float nan = qnan();
    float3 v00;
    bool ok00 = vertex_from_depth(L.depth, cols, u, v, fx_inv, fy_inv, cx, cy, v00);
    float* vm = L.vmap; float* nm = L.nmap;
    // Q7: like the reference, an invalid pixel only gets NaN in its x plane; the y/z planes keep what
    // they held (the integrate kernel can read a stale n_z for colour weighting, tsdf_volume.cu:601-622,
    // so reproducing the staleness keeps colour parity with the reference over a sequence).
    if (ok00) { vm[i] = v00.x; vm[i + P] = v00.y; vm[i + 2 * P] = v00.z; }
    else vm[i] = nan;
    bool okn = false;
    if (ok00 && u != cols - 1 && v != rows - 1) {
        float3 v01, v10;
        bool ok01 = vertex_from_depth(L.depth, cols, u + 1, v, fx_inv, fy_inv, cx, cy, v01);
        bool ok10 = vertex_from_depth(L.depth, cols, u, v + 1, fx_inv, fy_inv, cx, cy, v10);
        if (ok01 && ok10) {
            float3 n = normalized3(cross3(sub3(v01, v00), sub3(v10, v00)));
            nm[i] = n.x; nm[i + P] = n.y; nm[i + 2 * P] = n.z;
            okn = true;
        }
    }
    if (!okn) nm[i] = nan;
}

struct TransformParams { TransformLevel lv[LEVELS]; Mat33 R; float3 t; };

__global__ void __launch_bounds__(256)
transform_maps_kernel(const TransformParams p)
{
    const TransformLevel& L = p.lv[blockIdx.z];
    const int rows = L.rows, cols = L.cols;
    int x = threadIdx.x + blockIdx.x * blockDim.x;
    int y = threadIdx.y + blockIdx.y * blockDim.y;
    if (x >= cols || y >= rows) return;
    const float nan = qnan();
    const size_t P = (size_t)rows * cols, i = (size_t)y * cols + x;
    float3 vsrc, vdst = make_float3(nan, nan, nan);
    vsrc.x = L.vs[i];
    if (!isnan(vsrc.x)) {
        vsrc.y = L.vs[i + P]; vsrc.z = L.vs[i + 2 * P];
        vdst = add3(mul33(p.R, vsrc), p.t);
        L.vd[i + P] = vdst.y; L.vd[i + 2 * P] = vdst.z;
    }
    L.vd[i] = vdst.x;
    float3 nsrc, ndst = make_float3(nan, nan, nan);
    nsrc.x = L.ns[i];
    if (!isnan(nsrc.x)) {
        nsrc.y = L.ns[i + P]; nsrc.z = L.ns[i + 2 * P];
        ndst = mul33(p.R, nsrc);
        L.nd[i + P] = ndst.y; L.nd[i + 2 * P] = ndst.z;
    }
    L.nd[i] = ndst.x;
}

template <bool normalize>
__global__ void __launch_bounds__(256)
resize_map_kernel(int drows, int dcols, int srows, int scols, const float* __restrict__ input, float* __restrict__ output)
{
    int x = threadIdx.x + blockIdx.x * blockDim.x;
    int y = threadIdx.y + blockIdx.y * blockDim.y;
    if (x >= dcols || y >= drows) return;
    const float nan = qnan();
    const size_t SP = (size_t)srows * scols, DP = (size_t)drows * dcols;
    const size_t s = (size_t)(2 * y) * scols + 2 * x, d = (size_t)y * dcols + x;
    float x00 = input[s], x01 = input[s + 1], x10 = input[s + scols], x11 = input[s + scols + 1];
    if (isnan(x00) || isnan(x01) || isnan(x10) || isnan(x11)) {
        output[d] = nan;
        return;
    }
    float3 n;
    n.x = (x00 + x01 + x10 + x11) / 4;
    n.y = (input[s + SP] + input[s + SP + 1] + input[s + SP + scols] + input[s + SP + scols + 1]) / 4;
    n.z = (input[s + 2 * SP] + input[s + 2 * SP + 1] + input[s + 2 * SP + scols] + input[s + 2 * SP + scols + 1]) / 4;
    if (normalize) n = normalized3(n);
    output[d] = n.x; output[d + DP] = n.y; output[d + 2 * DP] = n.z;
}

} // namespace

int bilateral(const uint16_t* src, uint16_t* dst, int rows, int cols, cudaStream_t s)
{
    dim3 block(BIL_TX, BIL_TY), grid(div_up(cols, BIL_TX), div_up(rows, BIL_TY));
    bilateral_kernel<<<grid, block, 0, s>>>(src, dst, rows, cols, 0.5f / (SIGMA_SPACE * SIGMA_SPACE), 0.5f / (SIGMA_COLOR * SIGMA_COLOR));
    KT_LAUNCH_CHECK();
    return 0;
}

int pyrdown(const uint16_t* src, uint16_t* dst, int srows, int scols, cudaStream_t s)
{
    int drows = srows / 2, dcols = scols / 2;
    dim3 block(32, 8), grid(div_up(dcols, 32), div_up(drows, 8));
    pyrdown_gauss_kernel<<<grid, block, 0, s>>>(src, dst, srows, scols, drows, dcols, SIGMA_COLOR);
    KT_LAUNCH_CHECK();
    return 0;
}

int create_vmap(const Intr& k, const uint16_t* depth, float* vmap, int rows, int cols, cudaStream_t s)
{
    dim3 block(32, 8), grid(div_up(cols, 32), div_up(rows, 8));
    vmap_kernel<<<grid, block, 0, s>>>(depth, vmap, rows, cols, 1.f / k.fx, 1.f / k.fy, k.cx, k.cy);
    KT_LAUNCH_CHECK();
    return 0;
}

int create_nmap(const float* vmap, float* nmap, int rows, int cols, cudaStream_t s)
{
    dim3 block(32, 8), grid(div_up(cols, 32), div_up(rows, 8));
    nmap_kernel<<<grid, block, 0, s>>>(vmap, nmap, rows, cols);
    KT_LAUNCH_CHECK();
    return 0;
}

int create_maps_pyramid(const MapsLevel* levels, int n, cudaStream_t s)
{
    MapsParams p;
    for (int i = 0; i < n; ++i) { p.lv[i] = levels[i]; p.lv[i].fx_inv = 1.f / levels[i].k.fx; p.lv[i].fy_inv = 1.f / levels[i].k.fy; }
    for (int i = n; i < LEVELS; ++i) p.lv[i] = p.lv[n - 1];
    dim3 block(32, 8), grid(div_up(levels[0].cols, 32), div_up(levels[0].rows, 8), n);
    maps_pyramid_kernel<<<grid, block, 0, s>>>(p);
    KT_LAUNCH_CHECK();
    return 0;
}

int transform_maps_pyramid(const TransformLevel* levels, int n, const Mat33& R, const float3& t, cudaStream_t s)
{
    TransformParams p;
    for (int i = 0; i < n; ++i) p.lv[i] = levels[i];
    for (int i = n; i < LEVELS; ++i) p.lv[i] = levels[n - 1];
    p.R = R; p.t = t;
    dim3 block(32, 8), grid(div_up(levels[0].cols, 32), div_up(levels[0].rows, 8), n);
    transform_maps_kernel<<<grid, block, 0, s>>>(p);
    KT_LAUNCH_CHECK();
    return 0;
}

int transform_maps(const float* vs, const float* ns, const Mat33& R, const float3& t, float* vd, float* nd, int rows, int cols, cudaStream_t s)
{
    TransformLevel L = {vs, ns, vd, nd, rows, cols};
    return transform_maps_pyramid(&L, 1, R, t, s);
}

int resize_map(const float* in, float* out, int in_rows, int in_cols, bool normalize, cudaStream_t s)
{
    int drows = in_rows / 2, dcols = in_cols / 2;
    dim3 block(32, 8), grid(div_up(dcols, 32), div_up(drows, 8));
    if (normalize) resize_map_kernel<true><<<grid, block, 0, s>>>(drows, dcols, in_rows, in_cols, in, out);
    else resize_map_kernel<false><<<grid, block, 0, s>>>(drows, dcols, in_rows, in_cols, in, out);
    KT_LAUNCH_CHECK();
    return 0;
}

} // namespace kt
