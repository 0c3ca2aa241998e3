// kintinuous_b200 -- the GUI taps of the frontend: shaded / coloured model image and model depth rendered from the predicted surface.
//
// Reproduces (reference, src/frontend/cuda/image_generator.cu): generateImage / ImageGenerator (:56-186) and generateDepth /
// generateDepthKernel (:187-230), called from KintinuousTracker::getImage / getModelDepth / mutexOutLiveImage
// (KintinuousTracker.cpp:960-981, 1125-1154).  In the reference they are two launches + two cudaDeviceSynchronize inside processFrame
// whenever the GUI asks for an image; here ONE launch reads the model vertex / normal / colour maps once and writes whichever of the
// three outputs the caller wants, stream-ordered (no host sync: the tracker's tap copies the result out asynchronously).
// Per-pixel arithmetic is the reference's: heat-map colour of the voxel weight (channel w / 128), Lambert term |l . n| per light,
// + 20 offset with the r/b swap of :172-174, depth = row 3 of R^-1 applied to (v - t), * 1000, truncated to u16.
// Bound: HBM streaming (28 B read, 8 B written per pixel).
#include "kt_ops.h"

namespace kt {

namespace {

struct ViewParams {
    const float* vmap; const float* nmap; const uchar4* vcol; int rows, cols;
    float3 light[1]; int n_lights;
    uchar3* dst; uchar3* dst_color;                 // generateImage outputs (either may be null)
    float3 Rinv_row3, t; uint16_t* depth;           // generateDepth output (may be null)
};

__device__ __forceinline__ void heat_map_color(float value, int& red, int& green, int& blue)        // getHeatMapColor, image_generator.cu:73-101
{
    const float color[4][3] = {{0, 0, 1}, {0, 1, 0}, {1, 1, 0}, {1, 0, 0}};
    int idx1, idx2; float fractBetween = 0;
    if (value <= 0) idx1 = idx2 = 0;
    else if (value >= 1) idx1 = idx2 = 3;
    else { value = value * 3; idx1 = floor(value); idx2 = idx1 + 1; fractBetween = value - float(idx1); }
    red = ((color[idx2][0] - color[idx1][0]) * fractBetween + color[idx1][0]) * 235.0f;
    green = ((color[idx2][1] - color[idx1][1]) * fractBetween + color[idx1][1]) * 235.0f;
    blue = ((color[idx2][2] - color[idx1][2]) * fractBetween + color[idx1][2]) * 235.0f;
}

__global__ void __launch_bounds__(256)
views_kernel(const ViewParams p)
{
    const int x = threadIdx.x + blockIdx.x * blockDim.x, y = threadIdx.y + blockIdx.y * blockDim.y;
    if (x >= p.cols || y >= p.rows) return;
    const size_t P = (size_t)p.rows * p.cols, i = (size_t)y * p.cols + x;
    float3 v, n;
    v.x = p.vmap[i]; n.x = p.nmap[i];
    const bool ok = !isnan(v.x) && !isnan(n.x);
    uchar4 c4 = make_uchar4(0, 0, 0, 0);
    if (ok && (p.dst || p.dst_color)) c4 = p.vcol[i];
    if (ok) { v.y = p.vmap[i + P]; v.z = p.vmap[i + 2 * P]; }
    if (p.dst_color) {
        uchar3 color = make_uchar3(0, 0, 0);
        if (ok) color = make_uchar3(c4.x, c4.y, c4.z);            // (the reference clamps the already 8-bit channels to [0, 255])
        p.dst_color[i] = color;
    }
    if (p.dst) {
        uchar3 color = make_uchar3(0, 0, 0);
        if (ok) {
            n.y = p.nmap[i + P]; n.z = p.nmap[i + 2 * P];
            float weight = 1.f;
            for (int l = 0; l < p.n_lights; ++l) {
                const float3 vec = normalized3(sub3(p.light[l], v));
                weight *= fabs(dot3(vec, n));
            }
            int r, g, b;
            heat_map_color((float)c4.w / 128.0f, r, g, b);
            color = make_uchar3(b * weight + 20, g * weight + 20, r * weight + 20);
        }
        p.dst[i] = color;
    }
    if (p.depth) {
        unsigned short result = 0;
        if (ok) {
            const float v_z = dot3(p.Rinv_row3, sub3(v, p.t));
            result = static_cast<unsigned short>(v_z * 1000);
        }
        p.depth[i] = result;
    }
}

} // namespace

int generate_views(const float* vmap, const float* nmap, const uint8_t* vmap_color, int rows, int cols, const float* light_pos3, int n_lights,
                   uint8_t* dst_rgb, uint8_t* dst_color_rgb, const float* Rinv9, const float* t3, uint16_t* depth, cudaStream_t s)
{
    ViewParams p;
    p.vmap = vmap; p.nmap = nmap; p.vcol = (const uchar4*)vmap_color; p.rows = rows; p.cols = cols;
    p.n_lights = n_lights > 1 ? 1 : (n_lights < 0 ? 0 : n_lights);                 // LightSource holds one position (internal.h:289-293)
    p.light[0] = light_pos3 ? make_float3(light_pos3[0], light_pos3[1], light_pos3[2]) : make_float3(0, 0, 0);
    p.dst = (uchar3*)dst_rgb; p.dst_color = (uchar3*)dst_color_rgb; p.depth = depth;
    p.Rinv_row3 = Rinv9 ? make_float3(Rinv9[6], Rinv9[7], Rinv9[8]) : make_float3(0, 0, 1);
    p.t = t3 ? make_float3(t3[0], t3[1], t3[2]) : make_float3(0, 0, 0);
    dim3 block(32, 8), grid(div_up(cols, 32), div_up(rows, 8));
    views_kernel<<<grid, block, 0, s>>>(p);
    KT_LAUNCH_CHECK();
    return 0;
}

} // namespace kt
