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
        // ---- the operators really ran on the device: a flat wall 1.5 m in front of the camera ----
        std::vector<unsigned short> fh(rows * cols);
        f.download(&fh[0], cols * 2);
        int bad = 0;
        for (int i = 0; i < rows * cols; ++i) bad += (fh[i] != 1500);                    // bilateral of a constant image is the constant
        std::vector<float> vh(3 * rows * cols);
        vmap.download(&vh[0], cols * 4);
        for (int i = 0; i < rows * cols; ++i) { const float dz = vh[2 * rows * cols + i] - 1.5f; bad += (dz > 1e-6f || dz < -1e-6f); }  // vertex z = depth / 1000 (1500 * 0.001f)
        // integrate the wall into the 64^3 volume (camera at the volume centre, looking along +z), then extract and ray cast it
        std::vector<unsigned char> rgbh(rows * cols * 3, 200);
        DeviceArray2D<PixelRGB> colors; colors.upload(&rgbh[0], cols * 3, rows, cols);
        Mat33 I; I.data[0] = make_float3(1, 0, 0); I.data[1] = make_float3(0, 1, 0); I.data[2] = make_float3(0, 0, 1);
        const float3 tc = make_float3(1.5f, 1.5f, 0.2f);
        DeviceArray2D<float> scaled;
        Intr K(132.f, 132.f, 80.f, 66.75f);
        integrateTsdfVolume(PtrStepSz<unsigned short>(rows, cols, d.ptr(), d.step()), K, tsdf.getSize(), I, tc, tsdf.getTsdfTruncDist(),
                            PtrStep<short>(tsdf.data().ptr(), tsdf.data().step()), scaled, wrap, color.view(),
                            PtrStepSz<uchar3>(rows, cols, (uchar3*)colors.ptr(), colors.step()), nmap, true);
        DeviceArray<PointXYZRGB> wall = tsdf.fetchCloud(buf, wrap, color.view(), 0, 64, 0, 64, 0, 64, wrap);
        std::vector<PointXYZRGB> pts; wall.download(pts);
        int off = 0;
        for (size_t i = 0; i < pts.size(); ++i) { const float z = pts[i].z + 1.5f; if (z < 1.6f || z > 1.8f) ++off; }   // wall plane at 0.2 + 1.5 m (points are volume-centred)
        DeviceArray2D<float> rv(rows * 3, cols), rn(rows * 3, cols); DeviceArray2D<uchar4> rc(rows, cols);
        raycast(K, I, tc, tsdf.getTsdfTruncDist(), tsdf.getSize(), PtrStep<short>(tsdf.data().ptr(), tsdf.data().step()), rv, rn, wrap, rc, color.view());
        std::vector<float> rvh(3 * rows * cols); rv.download(&rvh[0], cols * 4);
        int hits = 0, wrong = 0;
        for (int i = 0; i < rows * cols; ++i) { const float z = rvh[2 * rows * cols + i]; if (z == z && rvh[i] == rvh[i]) { ++hits; if (z < 1.65f || z > 1.75f) ++wrong; } }
        std::printf("wall: %zu points (%d off-plane), raycast hits %d (%d off-plane), constant-image mismatches %d\n", pts.size(), off, hits, wrong, bad);
        if (bad || pts.size() < 500 || off * 50 > (int)pts.size() || hits < 1000 || wrong * 50 > hits) { std::printf("checks FAILED\n"); return 2; }
        std::printf("checks ok\n");
    } catch (const kt::Error& e) { std::printf("error: %s\n", e.what()); return 1; }
    return 0;
}
