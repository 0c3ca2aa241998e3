// kintinuous_b200_tracker.hpp -- the frontend's HOST CLASSES, source-compatible, over the C ABI (include/kintinuous_b200.h).
//
// A caller written against the reference's frontend -- backend/TrackerInterface.cpp:82-104 is the canonical one -- compiles against this
// header unchanged: same class names, member names, method names, argument lists and threading contract.
//   KintinuousTracker      KintinuousTracker.h:85-172   ctor from the 3x3 K, the 10-argument processFrame, finalise / reset /
//                                                         getVolumeOffset / getLastTranslation / getLastRotation / getVoxelSize /
//                                                         getCloudSlices / setOverlap / setParked / getLiveTsdf / getLiveImage,
//                                                         the public mutexes, condition variable, flags and dense pose graph
//   OdometryProvider       OdometryProvider.h:42-52     abstract provider; ICPOdometry (ICPOdometry.h:27-75) and RGBDOdometry
//                                                         (RGBDOdometry.h:37-110, with firstRun) implemented on the whole-frame kernels
//   CloudSlice             CloudSlice.h:28-129          the record handed to the backend
//   Resolution / Volume    Resolution.h:23-69, Volume.h:26-57
//   ThreadMutexObject<T>   utils/ThreadMutexObject.h
//
// The reference's interface types come from Eigen, OpenCV, Boost.Thread and PCL, none of which exists in this image.  namespace ktt
// ("tracker types") says where each one comes from:
//   default                      minimal stand-ins defined below (ktt::Vector3f with operator()(i), row-major ktt::Matrix3f, ktt::Matrix4f,
//                                ktt::Mat with at<double>(r, c), std::mutex / std::condition_variable_any, a points-only point cloud);
//   -DKT_TRACKER_USE_EIGEN_CV_BOOST_PCL   the real headers -- what a maintainer of the reference builds with.
// What the facade does NOT carry over: loadTrajectory (ground-truth odometry) and the place-recognition buffer are accepted and ignored
// (out of the hot path); the live image / live TSDF taps are served right after the frame was fused instead of
// right before (KintinuousTracker.cpp:835-862), i.e. they are one frame fresher.
#ifndef KINTINUOUS_B200_TRACKER_HPP_
#define KINTINUOUS_B200_TRACKER_HPP_

#include "kintinuous_b200_shim.hpp"
#include <cassert>
#include <cstring>
#include <string>
#include <vector>

#ifdef KT_TRACKER_USE_EIGEN_CV_BOOST_PCL
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <opencv2/core/core.hpp>
#include <boost/thread.hpp>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
namespace ktt {
typedef Eigen::Vector3f Vector3f;
typedef Eigen::Matrix<float, 3, 3, Eigen::RowMajor> Matrix3f;
typedef Eigen::Matrix4f Matrix4f;
typedef cv::Mat Mat;
typedef boost::mutex mutex;
typedef boost::condition_variable_any condition_variable_any;
typedef boost::mutex::scoped_lock scoped_lock;
typedef pcl::PointCloud<pcl::PointXYZRGB> PointCloud;
typedef pcl::PointCloud<pcl::PointXYZRGBNormal> PointCloudNormal;
}
#else
#include <condition_variable>
#include <mutex>
namespace ktt {
struct Vector3f { float v[3]; Vector3f() { v[0] = v[1] = v[2] = 0.f; } Vector3f(float a, float b, float c) { v[0] = a; v[1] = b; v[2] = c; }
    float& operator()(int i) { return v[i]; } const float& operator()(int i) const { return v[i]; } float* data() { return v; } const float* data() const { return v; } };
struct Matrix3f { float m[9]; Matrix3f() { for (int i = 0; i < 9; ++i) m[i] = (i % 4 == 0) ? 1.f : 0.f; }          // row-major, identity by default
    float& operator()(int r, int c) { return m[r * 3 + c]; } const float& operator()(int r, int c) const { return m[r * 3 + c]; } float* data() { return m; } const float* data() const { return m; } };
struct Matrix4f { float m[16]; Matrix4f() { for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.f : 0.f; }
    float& operator()(int r, int c) { return m[r * 4 + c]; } const float& operator()(int r, int c) const { return m[r * 4 + c]; } };
struct Mat { double d[9]; Mat() { for (int i = 0; i < 9; ++i) d[i] = 0; }                                               // the cv::Mat(3, 3, CV_64F) of the intrinsics
    template <class T> T& at(int r, int c) { return d[r * 3 + c]; } template <class T> const T& at(int r, int c) const { return d[r * 3 + c]; } };
typedef std::mutex mutex;
typedef std::condition_variable_any condition_variable_any;
struct scoped_lock { std::unique_lock<std::mutex> l; explicit scoped_lock(std::mutex& m) : l(m) {} void unlock() { l.unlock(); } void lock() { l.lock(); } };
struct PointCloud { std::vector<PointXYZRGB> points; };                                                                  // pcl::PointCloud<pcl::PointXYZRGB>::points
struct PointCloudNormal { std::vector<kt_point_xyzrgbnormal> points; };                                                  // pcl::PointCloud<pcl::PointXYZRGBNormal>::points (48-byte points)
}
#endif

// ---- Resolution.h:23-69, Volume.h:26-57 ----
class Resolution {
public:
    static const Resolution& get(int width = 0, int height = 0) { static const Resolution instance(width, height); return instance; }
    const int& width() const { return imgWidth; } const int& height() const { return imgHeight; }
    const int& cols() const { return imgWidth; } const int& rows() const { return imgHeight; } const int& numPixels() const { return imgNumPixels; }
private:
    Resolution(int w, int h) : imgWidth(w), imgHeight(h), imgNumPixels(w * h) { assert(w > 0 && h > 0); }
    const int imgWidth, imgHeight, imgNumPixels;
};
class Volume {
public:
    static Volume& get(float volumeSize = 0) { static Volume instance(volumeSize); return instance; }
    const float& getVolumeSize() { return volumeSize; }
private:
    explicit Volume(float s) : volumeSize(s) { assert(s > 0); }
    const float volumeSize;
};

// ---- utils/ThreadMutexObject.h ----
template <class T> class ThreadMutexObject {
public:
    ThreadMutexObject() {}
    ThreadMutexObject(T initialValue) : object(initialValue), lastCopy(initialValue) {}
    void assignValue(T newValue) { ktt::scoped_lock lock(mutex); object = lastCopy = newValue; }
    ktt::mutex& getMutex() { return mutex; }
    T& getReference() { return object; }
    void assignAndNotifyAll(T newValue) { ktt::scoped_lock lock(mutex); object = newValue; signal.notify_all(); }
    void notifyAll() { ktt::scoped_lock lock(mutex); signal.notify_all(); }
    T getValue() { ktt::scoped_lock lock(mutex); lastCopy = object; return lastCopy; }
    void operator++(int) { ktt::scoped_lock lock(mutex); object++; }
private:
    T object, lastCopy; ktt::mutex mutex; ktt::condition_variable_any signal;
};

// ---- the frontend options the reference reads from ConfigArgs::get() (ConfigArgs.h:113-169) ----
struct KtFrontendOptions {
    int vol;              // VOL (cuda/internal.h:243), a runtime value here
    int voxelShift;       // -t, 14
    bool useRGBD;         // -r
    bool useRGBDICP;      // -ri
    bool fastOdometry;    // -fo
    bool disableColorAngleWeight;   // -dc
    int gpu;              // -gpu
    // CloudSliceProcessor's work done on the device before a slice leaves it (backend/CloudSliceProcessor.cpp:97-162): when set, every
    // CloudSlice arrives with processedCloud filled (weight cull at weightCull = -cw, voxel grid, 20-NN normals) and the backend thread
    // has nothing left to do for it
    bool processSlicesOnGpu; int weightCull;
    const char* poseLog;  // "<saveFile>.poses" (outputPose, KintinuousTracker.cpp:199-218); NULL = off
    static KtFrontendOptions& get() { static KtFrontendOptions o = {512, 14, false, false, false, false, 0, false, 8, 0}; return o; }
};

// ---- CloudSlice.h:28-129 ----
class CloudSlice {
public:
    enum Dimension { XPlus, XMinus, YPlus, YMinus, ZPlus, ZMinus, FIRST, FINAL, TSDF };
    enum Odometry { ICP, GROUNDTRUTH, RGBD, FAIL };
    CloudSlice(ktt::PointCloud* cloud, Dimension dimension, Odometry odometry, ktt::Vector3f& cameraTranslation, ktt::Matrix3f& cameraRotation,
               uint64_t utime, uint64_t lagTime, unsigned char* rgbImage, unsigned char* tsdfImageColor = 0, unsigned char* tsdfImage = 0,
               unsigned short* depthData = 0, void* placeRecognitionFrame = 0)
        : cloud(cloud), processedCloud(0), dimension(dimension), odometry(odometry), cameraTranslation(cameraTranslation), cameraRotation(cameraRotation),
          utime(utime), lagTime(lagTime), rgbImage(0), tsdfImageColor(tsdfImageColor), tsdfImage(tsdfImage), depthData(0), placeRecognitionFrame(placeRecognitionFrame)
    {
        // CloudSlice.h:62-85: private copies of the frame's images
        if (rgbImage) { this->rgbImage = new unsigned char[Resolution::get().numPixels() * 3]; std::memcpy(this->rgbImage, rgbImage, Resolution::get().numPixels() * 3); }
        if (depthData) { this->depthData = new unsigned short[Resolution::get().numPixels()]; std::memcpy(this->depthData, depthData, Resolution::get().numPixels() * 2); }
    }
    virtual ~CloudSlice() { delete cloud; delete processedCloud; delete[] rgbImage; delete[] tsdfImageColor; delete[] tsdfImage; delete[] depthData; }
    ktt::PointCloud* cloud; ktt::PointCloudNormal* processedCloud;                    // CloudSlice.h:84-85
    Dimension dimension; Odometry odometry;
    ktt::Vector3f cameraTranslation; ktt::Matrix3f cameraRotation;
    uint64_t utime, lagTime;
    unsigned char* rgbImage; unsigned char* tsdfImageColor; unsigned char* tsdfImage; unsigned short* depthData; void* placeRecognitionFrame;
private:
    CloudSlice(const CloudSlice&); CloudSlice& operator=(const CloudSlice&);
};

// ---- OdometryProvider.h:42-52 and its two implementations on the whole-frame kernels ----
class OdometryProvider {
public:
    OdometryProvider() {}
    virtual ~OdometryProvider() {}
    virtual CloudSlice::Odometry getIncrementalTransformation(ktt::Vector3f& trans, ktt::Matrix3f& rot, const DeviceArray2D<unsigned short>& depth,
                                                              const DeviceArray2D<PixelRGB>& image, uint64_t timestamp, unsigned char* rgbImage, unsigned short* depthData) = 0;
    virtual void reset() = 0;
};

namespace kt { namespace shim {
// one odometry context (kernel scratch + photometric pyramids) per provider object; its own 32^3 volume is never used
inline kt_ctx* make_odometry_context(const Intr& intr, int odometry)
{
    kt_config cfg; std::memset(&cfg, 0, sizeof(cfg));
    cfg.rows = Resolution::get().rows(); cfg.cols = Resolution::get().cols(); cfg.fx = intr.fx; cfg.fy = intr.fy; cfg.cx = intr.cx; cfg.cy = intr.cy;
    cfg.vol = 32; cfg.volume_size = 6.f; cfg.odometry = odometry; cfg.fast_odometry = KtFrontendOptions::get().fastOdometry ? 1 : 0; cfg.voxel_shift = 14; cfg.overlap = 2;
    cfg.angle_color = 1; cfg.device = KtFrontendOptions::get().gpu; cfg.world = 1;
    kt_ctx* c = 0; kt::check(kt_create(&cfg, &c)); return c;
}
inline void level_pointers(const std::vector<DeviceArray2D<float> >& maps, const float* out[4]) { for (int l = 0; l < 4; ++l) out[l] = l < (int)maps.size() ? maps[l].ptr() : 0; }
}}

class ICPOdometry : public OdometryProvider {
public:
    ICPOdometry(std::vector<ktt::Vector3f>& tvecs_, std::vector<ktt::Matrix3f>& rmats_, std::vector<DeviceArray2D<float> >& vmaps_g_prev_,
                std::vector<DeviceArray2D<float> >& nmaps_g_prev_, std::vector<DeviceArray2D<float> >& vmaps_curr_, std::vector<DeviceArray2D<float> >& nmaps_curr_,
                Intr& intr, float /*distThresh*/ = 0.10f, float /*angleThresh*/ = 0.34202015f)
        : tvecs_(tvecs_), rmats_(rmats_), vmaps_g_prev_(vmaps_g_prev_), nmaps_g_prev_(nmaps_g_prev_), vmaps_curr_(vmaps_curr_), nmaps_curr_(nmaps_curr_),
          ctx_(kt::shim::make_odometry_context(intr, 0)) {}
    virtual ~ICPOdometry() { kt_destroy(ctx_); }
    // ICPOdometry.cpp:68-186 -- all levels and iterations in ONE cooperative launch, solve on the device
    CloudSlice::Odometry getIncrementalTransformation(ktt::Vector3f& trans, ktt::Matrix3f& rot, const DeviceArray2D<unsigned short>&, const DeviceArray2D<PixelRGB>&,
                                                      uint64_t, unsigned char*, unsigned short*)
    {
        const float* vg[4]; const float* ng[4]; const float* vc[4]; const float* nc[4];
        kt::shim::level_pointers(vmaps_g_prev_, vg); kt::shim::level_pointers(nmaps_g_prev_, ng); kt::shim::level_pointers(vmaps_curr_, vc); kt::shim::level_pointers(nmaps_curr_, nc);
        ktt::Matrix3f R; ktt::Vector3f t;
        kt::check(kt_odometry_increment(ctx_, 0, 0, rmats_.back().data(), tvecs_.back().data(), vg, ng, vc, nc, R.data(), t.data()));
        trans = t; rot = R;
        return CloudSlice::ICP;
    }
    void reset() {}
    static const int LEVELS = 4;
private:
    std::vector<ktt::Vector3f>& tvecs_; std::vector<ktt::Matrix3f>& rmats_;
    std::vector<DeviceArray2D<float> >& vmaps_g_prev_; std::vector<DeviceArray2D<float> >& nmaps_g_prev_;
    std::vector<DeviceArray2D<float> >& vmaps_curr_; std::vector<DeviceArray2D<float> >& nmaps_curr_;
    kt_ctx* ctx_;
};

class RGBDOdometry : public OdometryProvider {
public:
    RGBDOdometry(std::vector<ktt::Vector3f>& tvecs_, std::vector<ktt::Matrix3f>& rmats_, std::vector<DeviceArray2D<float> >& vmaps_g_prev_,
                 std::vector<DeviceArray2D<float> >& nmaps_g_prev_, std::vector<DeviceArray2D<float> >& vmaps_curr_, std::vector<DeviceArray2D<float> >& nmaps_curr_,
                 Intr& intr, float /*distThresh*/ = 0.10f, float /*angleThresh*/ = 0.34202015f)
        : tvecs_(tvecs_), rmats_(rmats_), vmaps_g_prev_(vmaps_g_prev_), nmaps_g_prev_(nmaps_g_prev_), vmaps_curr_(vmaps_curr_), nmaps_curr_(nmaps_curr_),
          ctx_(kt::shim::make_odometry_context(intr, KtFrontendOptions::get().useRGBDICP ? 2 : 1)) {}
    virtual ~RGBDOdometry() { kt_destroy(ctx_); }
    void firstRun(const DeviceArray2D<unsigned short>& depth, const DeviceArray2D<PixelRGB>& image)                 // RGBDOdometry.cpp:160-163
    { kt::check(kt_odometry_first_run(ctx_, depth.ptr(), (const uint8_t*)image.ptr())); }
    // RGBDOdometry.cpp:165-393 (photometric, or -ri: photometric + 100 x point-to-plane) in one cooperative launch
    CloudSlice::Odometry getIncrementalTransformation(ktt::Vector3f& trans, ktt::Matrix3f& rot, const DeviceArray2D<unsigned short>& depth, const DeviceArray2D<PixelRGB>& image,
                                                      uint64_t, unsigned char*, unsigned short*)
    {
        const float* vg[4]; const float* ng[4]; const float* vc[4]; const float* nc[4];
        kt::shim::level_pointers(vmaps_g_prev_, vg); kt::shim::level_pointers(nmaps_g_prev_, ng); kt::shim::level_pointers(vmaps_curr_, vc); kt::shim::level_pointers(nmaps_curr_, nc);
        const bool icp = KtFrontendOptions::get().useRGBDICP;
        ktt::Matrix3f R; ktt::Vector3f t;
        kt::check(kt_odometry_increment(ctx_, depth.ptr(), (const uint8_t*)image.ptr(), rmats_.back().data(), tvecs_.back().data(), vg, ng,
                                        icp ? vc : (const float* const*)0, icp ? nc : (const float* const*)0, R.data(), t.data()));
        trans = t; rot = R;
        return CloudSlice::RGBD;
    }
    void reset() {}
private:
    std::vector<ktt::Vector3f>& tvecs_; std::vector<ktt::Matrix3f>& rmats_;
    std::vector<DeviceArray2D<float> >& vmaps_g_prev_; std::vector<DeviceArray2D<float> >& nmaps_g_prev_;
    std::vector<DeviceArray2D<float> >& vmaps_curr_; std::vector<DeviceArray2D<float> >& nmaps_curr_;
    kt_ctx* ctx_;
};

// ---- KintinuousTracker.h:85-172 ----
class KintinuousTracker {
public:
    ThreadMutexObject<bool> tsdfRequest;
    bool tsdfAvailable;
    ktt::mutex tsdfMutex;
    bool imageAvailable;
    ktt::mutex imageMutex;
    bool cycledMutex;
    ktt::mutex cloudMutex;
    ktt::condition_variable_any cloudSignal;

    // depthIntrinsics: 3x3, CV_64F (MainController.cpp:222-227); everything else comes from Resolution / Volume / KtFrontendOptions
    explicit KintinuousTracker(ktt::Mat* depthIntrinsics)
        : tsdfRequest(false), tsdfAvailable(false), imageAvailable(false), cycledMutex(false), init_utime(0), firstRgbImage(0), firstDepthData(0),
          lastRgbImage(0), lastDepthData(0), lastOdometry(CloudSlice::ICP), placeRecognitionId(0), latestDensePoseId(0), ctx_(0), handed_(0), liveTsdf(0), liveImage(0)
    {
        const KtFrontendOptions& o = KtFrontendOptions::get();
        kt_config cfg; std::memset(&cfg, 0, sizeof(cfg));
        cfg.rows = Resolution::get().rows(); cfg.cols = Resolution::get().cols();
        cfg.fx = (float)depthIntrinsics->at<double>(0, 0); cfg.fy = (float)depthIntrinsics->at<double>(1, 1);              // KintinuousTracker.cpp:86-89
        cfg.cx = (float)depthIntrinsics->at<double>(0, 2); cfg.cy = (float)depthIntrinsics->at<double>(1, 2);
        cfg.vol = o.vol; cfg.volume_size = Volume::get().getVolumeSize();
        cfg.odometry = o.useRGBDICP ? 2 : (o.useRGBD ? 1 : 0); cfg.fast_odometry = o.fastOdometry ? 1 : 0;
        cfg.voxel_shift = o.voxelShift; cfg.overlap = 2; cfg.angle_color = o.disableColorAngleWeight ? 0 : 1; cfg.device = o.gpu; cfg.world = 1;
        kt::check(kt_create(&cfg, &ctx_));
        if (o.processSlicesOnGpu) kt::check(kt_set_slice_processing(ctx_, 1, o.weightCull));
        if (o.poseLog) kt::check(kt_set_pose_log(ctx_, o.poseLog));
        lastOdometry = cfg.odometry == 0 ? CloudSlice::ICP : CloudSlice::RGBD;
    }
    virtual ~KintinuousTracker()
    {
        for (size_t i = 0; i < sharedCloudSlices.size(); ++i) delete sharedCloudSlices[i];
        delete liveTsdf; delete liveImage;
        kt_destroy(ctx_);
    }

    void processFrame(const DeviceArray2D<unsigned short>& depth, const DeviceArray2D<PixelRGB>& colors, unsigned char* rgbImage, unsigned short* depthData,
                      uint64_t timestamp, bool /*compression*/, uint8_t* /*lastCompressedDepth*/, int /*depthSize*/, uint8_t* /*lastCompressedImage*/, int /*imageSize*/)
    {
        lastRgbImage = rgbImage; lastDepthData = depthData;
        kt_pose p;
        kt::check(kt_process_frame_device(ctx_, depth.ptr(), (const uint8_t*)colors.ptr(), timestamp, &p));
        if (p.frame == 1) { init_utime.assignValue(timestamp); firstRgbImage.assignValue(rgbImage); firstDepthData.assignValue(depthData); }     // .cpp:497-503
        // mutexOutCloudBuffer (.cpp:1156-1208): every slab that left the volume during this frame is handed to the backend
        while (handed_ < kt_num_slices(ctx_)) handOver(handed_++);
        // dense pose graph (.cpp:529-536, :901-909): mirrored from the context (the first frame carries the loop-pose flag)
        while ((int)densePoseGraph.size() < kt_num_dense_poses(ctx_)) {
            kt_dense_pose d; kt::check(kt_get_dense_pose(ctx_, (int)densePoseGraph.size(), &d));
            ktt::Matrix4f pose;
            for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) pose(r, c) = d.pose[r * 4 + c];
            densePoseGraph.push_back(DensePose(d.timestamp, pose, d.is_loop_pose != 0));
            latestDensePoseId++;
        }
        // GUI taps (.cpp:835-862)
        if (tsdfRequest.getValue()) {
            bool needed; { ktt::scoped_lock l(tsdfMutex); needed = !tsdfAvailable; }
            if (needed) mutexOutLiveTsdf(p, timestamp);
        }
        bool imageNeeded; { ktt::scoped_lock l(imageMutex); imageNeeded = !imageAvailable; }
        if (imageNeeded) mutexOutLiveImage(p, timestamp);
        placeRecognitionId++;
    }

    ktt::Vector3f getVolumeOffset() const { const float h = Volume::get().getVolumeSize() * 0.5f; return ktt::Vector3f(h, h, h); }            // volumeBasis
    void setParked(const bool park) { kt::check(kt_set_parked(ctx_, park ? 1 : 0)); }
    ktt::Vector3f getLastTranslation() const { kt_pose p; kt::check(kt_get_pose(ctx_, &p)); const ktt::Vector3f b = getVolumeOffset(); return ktt::Vector3f(p.t[0] - b(0), p.t[1] - b(1), p.t[2] - b(2)); }
    ktt::Vector3f getVoxelSize() const { const float v = kt_get_voxel_size(ctx_); return ktt::Vector3f(v, v, v); }
    void finalise() { kt::check(kt_finalise(ctx_)); while (handed_ < kt_num_slices(ctx_)) handOver(handed_++); }                               // .cpp:1003-1048
    ktt::Matrix3f getLastRotation() const { kt_pose p; kt::check(kt_get_pose(ctx_, &p)); ktt::Matrix3f R; for (int i = 0; i < 9; ++i) R.data()[i] = p.R[i]; return R; }
    std::vector<CloudSlice*>& getCloudSlices() { return sharedCloudSlices; }
    void setOverlap(int overlap) { kt::check(kt_set_overlap(ctx_, overlap)); }
    CloudSlice* getLiveTsdf() { return liveTsdf; }
    CloudSlice* getLiveImage() { return liveImage; }
    void reset()
    {
        kt::check(kt_reset(ctx_));
        for (size_t i = 0; i < sharedCloudSlices.size(); ++i) delete sharedCloudSlices[i];
        sharedCloudSlices.clear(); handed_ = 0; densePoseGraph.clear(); latestDensePoseId.assignValue(0);
    }
    void loadTrajectory(const std::string&) {}                    // ground-truth odometry: not on the hot path

    ThreadMutexObject<uint64_t> init_utime;
    ThreadMutexObject<unsigned char*> firstRgbImage;
    ThreadMutexObject<unsigned short*> firstDepthData;
    unsigned char* lastRgbImage;
    unsigned short* lastDepthData;
    CloudSlice::Odometry lastOdometry;
    ThreadMutexObject<int> placeRecognitionId;

    class DensePose {
    public:
        DensePose(uint64_t timestamp, ktt::Matrix4f pose, bool isLoopPose) : timestamp(timestamp), pose(pose), isLoopPose(isLoopPose) {}
        DensePose() {}
        uint64_t timestamp; ktt::Matrix4f pose; bool isLoopPose;
    };
    std::vector<DensePose> densePoseGraph;
    ThreadMutexObject<int> latestDensePoseId;

    kt_ctx* handle() { return ctx_; }

private:
    void handOver(int i)
    {
        kt_slice_info info; kt::check(kt_get_slice_info(ctx_, i, &info));
        ktt::PointCloud* cloud = new ktt::PointCloud();
        cloud->points.resize(info.count);
        size_t n = info.count;
        if (n) kt::check(kt_get_slice(ctx_, i, (kt_point_xyzrgb*)&cloud->points[0], n, &n, 0, 0));
        ktt::Vector3f t(info.camera_t[0], info.camera_t[1], info.camera_t[2]);
        ktt::Matrix3f R; for (int k = 0; k < 9; ++k) R.data()[k] = info.camera_R[k];
        CloudSlice* s = new CloudSlice(cloud, (CloudSlice::Dimension)info.dimension, (CloudSlice::Odometry)info.odometry, t, R, info.utime, 0,
                                       info.dimension == CloudSlice::FINAL ? lastRgbImage : 0);
        if (KtFrontendOptions::get().processSlicesOnGpu) {
            size_t m = 0; kt::check(kt_get_processed_slice(ctx_, i, 0, 0, &m));
            s->processedCloud = new ktt::PointCloudNormal(); s->processedCloud->points.resize(m);
            if (m) kt::check(kt_get_processed_slice(ctx_, i, (kt_point_xyzrgbnormal*)&s->processedCloud->points[0], m, &m));
        }
        { ktt::scoped_lock lock(cloudMutex); cycledMutex = true; sharedCloudSlices.push_back(s); }
        cloudSignal.notify_all();
    }
    void mutexOutLiveTsdf(const kt_pose& p, uint64_t utime)       // .cpp:1087-1123
    {
        size_t n = 0; const size_t cap = (size_t)3 * Resolution::get().numPixels();          // the tracker's cloud buffer (KintinuousTracker.cpp:77)
        ktt::PointCloud* cloud = new ktt::PointCloud(); cloud->points.resize(cap);
        kt::check(kt_get_live_tsdf(ctx_, (kt_point_xyzrgb*)&cloud->points[0], cap, &n));
        cloud->points.resize(n < cap ? n : cap);
        ktt::Vector3f t(p.global_t[0], p.global_t[1], p.global_t[2]); ktt::Matrix3f R; for (int k = 0; k < 9; ++k) R.data()[k] = p.R[k];
        ktt::scoped_lock l(tsdfMutex);
        tsdfAvailable = true;
        delete liveTsdf;
        liveTsdf = new CloudSlice(cloud, CloudSlice::TSDF, lastOdometry, t, R, utime, 0, 0);
    }
    void mutexOutLiveImage(const kt_pose& p, uint64_t utime)      // .cpp:1125-1154
    {
        const int np = Resolution::get().numPixels();
        unsigned char* tsdfImageColor = new unsigned char[np * 3]; unsigned char* tsdfImage = new unsigned char[np * 3];
        kt::check(kt_get_live_image(ctx_, tsdfImage, tsdfImageColor, 0));
        ktt::Vector3f t(p.global_t[0], p.global_t[1], p.global_t[2]); ktt::Matrix3f R; for (int k = 0; k < 9; ++k) R.data()[k] = p.R[k];
        ktt::scoped_lock l(imageMutex);
        imageAvailable = true;
        delete liveImage;
        liveImage = new CloudSlice(0, CloudSlice::TSDF, lastOdometry, t, R, utime, 0, lastRgbImage, tsdfImageColor, tsdfImage, lastDepthData);
    }
    KintinuousTracker(const KintinuousTracker&); KintinuousTracker& operator=(const KintinuousTracker&);
    kt_ctx* ctx_; int handed_;
    std::vector<CloudSlice*> sharedCloudSlices;
    CloudSlice* liveTsdf; CloudSlice* liveImage;
};

#endif // KINTINUOUS_B200_TRACKER_HPP_
