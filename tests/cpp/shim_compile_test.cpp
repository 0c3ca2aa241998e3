// Compile-only check of the C++ shim: a caller written against the reference's operator names (cuda/internal.h) builds
// against include/kintinuous_b200_shim.hpp and links against libkintinuous_b200.so.  Run (no GPU needed to build):
//   g++ -std=c++14 -I include -I /usr/local/cuda/include tests/cpp/shim_compile_test.cpp -L kintinuous_b200 -lkintinuous_b200 -L /usr/local/cuda/lib64 -lcudart
#include "kintinuous_b200_shim.hpp"
#include <cstdio>

int main()
{
    if (!kt_cuda_available()) { std::printf("no CUDA device: shim links, nothing to run\n"); return 0; }
    try {
        const int rows = 120, cols = 160;
        std::vector<unsigned short> depth(rows * cols, 1500);
        DeviceArray2D<unsigned short> d, f, p;
        d.upload(&depth[0], cols * 2, rows, cols);
        bilateralFilter(d, f);
        pyrDown(f, p);
        DeviceArray2D<float> vmap, nmap;
        createVMap(Intr(132.f, 132.f, 80.f, 66.75f), f, vmap);
        createNMap(vmap, nmap);
        std::printf("shim ok: %d x %d -> %d x %d\n", f.rows(), f.cols(), p.rows(), p.cols());
        // host classes of the volume (TSDFVolume.h / ColorVolume.h) on top of the operator API
        TsdfVolume tsdf(64, 3.f);
        ColorVolume color(tsdf);
        tsdf.reset(color);
        tsdf.setTsdfTruncDist(0.01f);                       // clamped to 2.1 voxels like the reference
        DeviceArray<PointXYZRGB> buf; int3 wrap = make_int3(0, 0, 0);
        DeviceArray<PointXYZRGB> got = tsdf.fetchCloud(buf, wrap, color.view(), 0, 64, 0, 64, 0, 64, wrap);
        std::vector<float> v; std::vector<short> w;
        tsdf.downloadTsdfAndWeighs(color, v, w);
        std::printf("volume ok: trunc %.4f, %zu points, %zu voxels\n", tsdf.getTsdfTruncDist(), got.size(), v.size());
    } catch (const kt::Error& e) { std::printf("error: %s\n", e.what()); return 1; }
    return 0;
}
