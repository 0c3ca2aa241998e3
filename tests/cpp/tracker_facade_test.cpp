// A caller written against the reference frontend's HOST CLASSES -- the calls backend/TrackerInterface.cpp:82-104 makes per frame (host
// frame -> PtrStepSz views -> DeviceArray2D::upload -> the 10-argument KintinuousTracker::processFrame), MainController's construction
// (Resolution / Volume singletons, 3x3 K), the backend's consumption of getCloudSlices() under cloudMutex, the GUI's live image tap and an
// OdometryProvider used on its own -- compiled against include/kintinuous_b200_tracker.hpp and run on the device.
//   g++ -std=c++14 -I include -I /usr/local/cuda/include tests/cpp/tracker_facade_test.cpp -L kintinuous_b200 -lkintinuous_b200 -L /usr/local/cuda/lib64 -lcudart -lpthread
#include "kintinuous_b200_tracker.hpp"
#include <cmath>
#include <cstdio>

// a log reader with the members TrackerInterface reads (utils/RawLogReader.h)
struct LogReaderStub {
    unsigned char* decompressedImage; unsigned short* decompressedDepth; int64_t timestamp; bool isCompressed;
    unsigned char* compressedDepth; int compressedDepthSize; unsigned char* compressedImage; int compressedImageSize;
};

// synthetic scene: the inside of a 2.4 x 1.8 x 5 m corridor (side walls, floor and ceiling in view: all six degrees of freedom are
// observable) seen from a camera that moves 4 cm per frame along +x
static void render(int k, int rows, int cols, float fx, float fy, float cx, float cy, std::vector<unsigned short>& depth, std::vector<unsigned char>& rgb)
{
    const float ox = 0.04f * k, half[3] = {1.2f, 0.9f, 2.5f};
    for (int v = 0; v < rows; ++v)
        for (int u = 0; u < cols; ++u) {
            const float d[3] = {(u - cx) / fx, (v - cy) / fy, 1.f}, o[3] = {ox, 0.f, 0.f};
            float t = 1e30f;
            for (int a = 0; a < 3; ++a) if (d[a] != 0.f) { const float tt = ((d[a] > 0 ? half[a] : -half[a]) - o[a]) / d[a]; if (tt < t) t = tt; }
            depth[v * cols + u] = (unsigned short)std::lround(1000.0 * t);
            const float px = o[0] + d[0] * t, py = d[1] * t, pz = d[2] * t;
            for (int ch = 0; ch < 3; ++ch) rgb[(v * cols + u) * 3 + ch] = (unsigned char)(128 + 100 * std::sin(7 * px + ch) * std::sin(5 * py + 2 * ch) * std::sin(6 * pz - ch));
        }
}

int main()
{
    if (!kt_cuda_available()) { std::printf("no CUDA device: facade links, nothing to run\n"); return 0; }
    try {
        const int cols = 160, rows = 120;
        Resolution::get(cols, rows);
        Volume::get(6.0f);
        KtFrontendOptions::get().vol = 128; KtFrontendOptions::get().voxelShift = 2;
        KtFrontendOptions::get().processSlicesOnGpu = true;          // CloudSlice::processedCloud filled on the device (CloudSliceProcessor.cpp:97-162)
        ktt::Mat K;
        K.at<double>(0, 0) = 132.0; K.at<double>(1, 1) = 132.0; K.at<double>(0, 2) = 80.0; K.at<double>(1, 2) = 66.75; K.at<double>(2, 2) = 1.0;
        KintinuousTracker* frontend = new KintinuousTracker(&K);

        std::vector<unsigned short> depthHost(rows * cols); std::vector<unsigned char> rgbHost(rows * cols * 3);
        LogReaderStub reader = {&rgbHost[0], &depthHost[0], 0, false, 0, 0, 0, 0};
        LogReaderStub* logRead = &reader;
        PtrStepSz<const unsigned short> depth; PtrStepSz<const PixelRGB> rgb24;
        DeviceArray2D<unsigned short> depth_device; DeviceArray2D<PixelRGB> colors_device;

        for (int currentFrame = 0; currentFrame < 14; ++currentFrame) {
            render(currentFrame, rows, cols, 132.f, 132.f, 80.f, 66.75f, depthHost, rgbHost);
            logRead->timestamp = 1000 + 33 * currentFrame;
            // ---- the per-frame body of TrackerInterface::process ----
            depth.data = (unsigned short*)logRead->decompressedDepth;
            rgb24.data = (PixelRGB*)logRead->decompressedImage;
            depth.step = Resolution::get().width() * 2; depth.rows = Resolution::get().rows(); depth.cols = Resolution::get().cols();
            rgb24.step = Resolution::get().width() * 3; rgb24.rows = Resolution::get().rows(); rgb24.cols = Resolution::get().cols();
            depth_device.upload(depth.data, depth.step, depth.rows, depth.cols);
            colors_device.upload(rgb24.data, rgb24.step, rgb24.rows, rgb24.cols);
            frontend->processFrame(depth_device, colors_device, logRead->decompressedImage, logRead->decompressedDepth, logRead->timestamp, logRead->isCompressed,
                                   logRead->compressedDepth, logRead->compressedDepthSize, logRead->compressedImage, logRead->compressedImageSize);
        }
        // the tracker followed the camera: 13 frames x 4 cm along x
        ktt::Vector3f t = frontend->getLastTranslation();
        ktt::Matrix3f R = frontend->getLastRotation();
        std::printf("translation %.4f %.4f %.4f  R00 %.5f  voxel %.4f  poses %zu\n", t(0), t(1), t(2), R(0, 0), frontend->getVoxelSize()(0), frontend->densePoseGraph.size());
        int bad = 0;
        // the volume shifted along +x (2-voxel threshold at 4.7 cm voxels): the translation relative to the volume centre wraps, the dense pose graph keeps the global one
        const float gx = frontend->densePoseGraph.back().pose(0, 3);
        if (std::fabs(gx - 0.52f) > 0.03f || std::fabs(frontend->densePoseGraph.back().pose(1, 3)) > 0.02f) { std::printf("global pose off: %.4f\n", gx); ++bad; }
        if (frontend->densePoseGraph.size() != 14 || frontend->latestDensePoseId.getValue() != 14 || frontend->init_utime.getValue() != 1000) ++bad;
        // backend side: consume the slices under the mutex like CloudSliceProcessor does
        size_t npts = 0; int nslices = 0;
        {
            ktt::scoped_lock lock(frontend->cloudMutex);
            std::vector<CloudSlice*>& slices = frontend->getCloudSlices();
            nslices = (int)slices.size();
            for (size_t i = 0; i < slices.size(); ++i) { npts += slices[i]->cloud->points.size(); if (slices[i]->dimension != CloudSlice::XPlus || slices[i]->odometry != CloudSlice::ICP) ++bad; }
        }
        if (nslices < 3 || !frontend->cycledMutex) { std::printf("expected +x shifts, got %d slices\n", nslices); ++bad; }
        // GUI side: the live image was produced on the first frame (imageAvailable was false) and stays until the GUI clears the flag
        CloudSlice* live = frontend->getLiveImage();
        if (!live || !live->tsdfImage || !live->tsdfImageColor || !frontend->imageAvailable) { std::printf("no live image\n"); ++bad; }
        { ktt::scoped_lock l(frontend->imageMutex); frontend->imageAvailable = false; }
        frontend->tsdfRequest.assignValue(true);
        render(14, rows, cols, 132.f, 132.f, 80.f, 66.75f, depthHost, rgbHost);
        depth_device.upload(&depthHost[0], cols * 2, rows, cols); colors_device.upload(&rgbHost[0], cols * 3, rows, cols);
        frontend->processFrame(depth_device, colors_device, &rgbHost[0], &depthHost[0], 2000, false, 0, 0, 0, 0);
        live = frontend->getLiveImage();
        int lit = 0; for (int i = 0; live && i < rows * cols; ++i) lit += (live->tsdfImage[i * 3] | live->tsdfImage[i * 3 + 1] | live->tsdfImage[i * 3 + 2]) != 0;
        CloudSlice* ltsdf = frontend->getLiveTsdf();
        std::printf("slices %d (%zu points), live image lit pixels %d, live tsdf points %zu\n", nslices, npts, lit, ltsdf ? ltsdf->cloud->points.size() : 0);
        if (lit < rows * cols / 2 || !ltsdf || ltsdf->cloud->points.size() < 1000 || ltsdf->dimension != CloudSlice::TSDF) ++bad;
        frontend->finalise();
        if (frontend->getCloudSlices().back()->dimension != CloudSlice::FINAL) ++bad;
        {
            CloudSlice* fin = frontend->getCloudSlices().back();
            const size_t np = fin->processedCloud ? fin->processedCloud->points.size() : 0;
            float nn = 0.f;
            if (np) { const kt_point_xyzrgbnormal& q = fin->processedCloud->points[np / 2]; nn = q.nx * q.nx + q.ny * q.ny + q.nz * q.nz; }
            std::printf("FINAL slice: %zu points, processedCloud %zu points, |n|^2 of one = %.4f\n", fin->cloud->points.size(), np, nn);
            if (np < 100 || np > fin->cloud->points.size() || std::fabs(nn - 1.f) > 1e-3f) ++bad;
        }

        // ---- an OdometryProvider on its own: ICPOdometry over maps the caller owns (what KintinuousTracker.cpp:562-577 does) ----
        {
            render(0, rows, cols, 132.f, 132.f, 80.f, 66.75f, depthHost, rgbHost);
            DeviceArray2D<unsigned short> d0, f0; d0.upload(&depthHost[0], cols * 2, rows, cols);
            std::vector<DeviceArray2D<float> > vg(4), ng(4), vc(4), nc(4);
            std::vector<DeviceArray2D<unsigned short> > pyr(4);
            Intr intr(132.f, 132.f, 80.f, 66.75f);
            Mat33 Rm; Rm.data[0] = make_float3(1, 0, 0); Rm.data[1] = make_float3(0, 1, 0); Rm.data[2] = make_float3(0, 0, 1);
            const float3 tm = make_float3(3.f, 3.f, 3.f);
            bilateralFilter(d0, pyr[0]);
            for (int l = 1; l < 4; ++l) pyrDown(pyr[l - 1], pyr[l]);
            for (int l = 0; l < 4; ++l) { DeviceArray2D<float> v, n; createVMap(intr(l), pyr[l], v); createNMap(v, n); tranformMaps(v, n, Rm, tm, vg[l], ng[l]); }     // model = frame 0 in the volume frame
            render(2, rows, cols, 132.f, 132.f, 80.f, 66.75f, depthHost, rgbHost);                                                                                // current = 8 cm further along x
            d0.upload(&depthHost[0], cols * 2, rows, cols);
            bilateralFilter(d0, pyr[0]);
            for (int l = 1; l < 4; ++l) pyrDown(pyr[l - 1], pyr[l]);
            for (int l = 0; l < 4; ++l) { createVMap(intr(l), pyr[l], vc[l]); createNMap(vc[l], nc[l]); }
            std::vector<ktt::Vector3f> tvecs(1, ktt::Vector3f(3.f, 3.f, 3.f)); std::vector<ktt::Matrix3f> rmats(1);
            ICPOdometry icp(tvecs, rmats, vg, ng, vc, nc, intr);
            OdometryProvider* odom = &icp;
            ktt::Vector3f tr; ktt::Matrix3f rot;
            CloudSlice::Odometry kind = odom->getIncrementalTransformation(tr, rot, d0, colors_device, 0, 0, 0);
            std::printf("ICPOdometry: t = %.4f %.4f %.4f (expected 3.08 3 3)\n", tr(0), tr(1), tr(2));
            if (kind != CloudSlice::ICP || std::fabs(tr(0) - 3.08f) > 0.01f || std::fabs(tr(1) - 3.f) > 0.01f || std::fabs(tr(2) - 3.f) > 0.01f || std::fabs(rot(0, 0) - 1.f) > 1e-3f) ++bad;
        }
        delete frontend;
        if (bad) { std::printf("facade checks FAILED (%d)\n", bad); return 2; }
        std::printf("facade checks ok\n");
    } catch (const kt::Error& e) { std::printf("error: %s\n", e.what()); return 1; }
    return 0;
}
