// kintinuous_b200_shim.hpp -- source-compatible C++ shim over the C ABI (include/kintinuous_b200.h).
//
// Re-creates, with the SAME names, argument order and meaning, the part of the reference's frontend that the hot path
// exposes, so that the reference's host classes (or a maintainer's fork of them) can call this library instead of
// src/frontend/cuda/*.cu:
//   * containers: DeviceArray<T>, DeviceArray2D<T>, PtrStep / PtrStepSz / PtrSz views   (cuda/containers/*.hpp)
//   * POD types:  Intr, IntrDoublePrecision, Mat33, PixelRGB, PointXYZRGB, JtJJtrSE3, DataTerm  (cuda/internal.h:90-292)
//   * the operator API: every free function of cuda/internal.h:299-536
//   * class TsdfVolume (TSDFVolume.h:68-158) and a KintinuousTracker facade (KintinuousTracker.h:85-172) without the
//     Eigen / OpenCV / PCL / Boost types (those are absent from this image; INTEGRATION.md shows the 10-line adapters).
// Differences by design: device images are allocated with a COMPACT pitch (step == cols * sizeof(T)); errors throw
// kt::Error instead of exit(0) (cuda/internal.h:76-86); VOL is a runtime value (kt::shim::set_volume_resolution).
#ifndef KINTINUOUS_B200_SHIM_HPP_
#define KINTINUOUS_B200_SHIM_HPP_

#include <cuda_runtime_api.h>
#include <vector_types.h>
#include <vector_functions.h>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>
#include "kintinuous_b200.h"

namespace kt {
struct Error : std::runtime_error { explicit Error(const std::string& m) : std::runtime_error(m) {} };
inline void check(int status) { if (status != KT_OK) throw Error(std::string("kintinuous_b200: ") + kt_last_error()); }
inline void cuda(cudaError_t e) { if (e != cudaSuccess) throw Error(std::string("CUDA: ") + cudaGetErrorString(e)); }
namespace shim {
inline int& volume_resolution() { static int v = 512; return v; }          // '#define VOL 512' (cuda/internal.h:243)
inline void set_volume_resolution(int v) { volume_resolution() = v; }
}
}

// ---- kernel views (cuda/containers/kernel_containers.hpp:49-93) ----
template <typename T> struct DevPtr { T* data; DevPtr() : data(0) {} DevPtr(T* d) : data(d) {} operator T*() { return data; } operator const T*() const { return data; } };
template <typename T> struct PtrSz : DevPtr<T> { size_t size; PtrSz() : size(0) {} PtrSz(T* d, size_t s) : DevPtr<T>(d), size(s) {} };
template <typename T> struct PtrStep : DevPtr<T> { size_t step; PtrStep() : step(0) {} PtrStep(T* d, size_t s) : DevPtr<T>(d), step(s) {}
    T* ptr(int y = 0) { return (T*)((char*)DevPtr<T>::data + y * step); } const T* ptr(int y = 0) const { return (const T*)((const char*)DevPtr<T>::data + y * step); } };
template <typename T> struct PtrStepSz : PtrStep<T> { int cols, rows; PtrStepSz() : cols(0), rows(0) {} PtrStepSz(int r, int c, T* d, size_t s) : PtrStep<T>(d, s), cols(c), rows(r) {} };

// ---- containers (cuda/containers/device_array.hpp) : refcount-free RAII, compact pitch ----
template <class T> class DeviceArray {
public:
    DeviceArray() : data_(0), size_(0), own_(false) {}
    explicit DeviceArray(size_t n) : data_(0), size_(0), own_(false) { create(n); }
    DeviceArray(T* p, size_t n) : data_(p), size_(n), own_(false) {}
    DeviceArray(const DeviceArray& o) : data_(o.data_), size_(o.size_), own_(false) {}          // views, like the reference's refcounted copies
    DeviceArray& operator=(const DeviceArray& o) { if (this != &o) { release(); data_ = o.data_; size_ = o.size_; own_ = false; } return *this; }
    ~DeviceArray() { release(); }
    void create(size_t n) { if (n == size_) return; release(); kt::cuda(cudaMalloc((void**)&data_, n * sizeof(T))); size_ = n; own_ = true; }
    void release() { if (own_ && data_) cudaFree(data_); data_ = 0; size_ = 0; own_ = false; }
    void upload(const T* h, size_t n) { create(n); kt::cuda(cudaMemcpy(data_, h, n * sizeof(T), cudaMemcpyHostToDevice)); }
    void download(T* h) const { kt::cuda(cudaMemcpy(h, data_, size_ * sizeof(T), cudaMemcpyDeviceToHost)); }
    template <class A> void download(std::vector<T, A>& v) const { v.resize(size_); if (size_) download(&v[0]); }
    T* ptr() { return data_; } const T* ptr() const { return data_; }
    size_t size() const { return size_; } bool empty() const { return !data_; }
    operator T*() { return data_; } operator const T*() const { return data_; }
    operator PtrSz<T>() const { return PtrSz<T>(data_, size_); }
private:
    T* data_; size_t size_; bool own_;
};

template <class T> class DeviceArray2D {
public:
    DeviceArray2D() : data_(0), rows_(0), cols_(0), own_(false) {}
    DeviceArray2D(int r, int c) : data_(0), rows_(0), cols_(0), own_(false) { create(r, c); }
    DeviceArray2D(int r, int c, void* d, size_t /*stepBytes*/) : data_((T*)d), rows_(r), cols_(c), own_(false) {}
    DeviceArray2D(const DeviceArray2D& o) : data_(o.data_), rows_(o.rows_), cols_(o.cols_), own_(false) {}
    DeviceArray2D& operator=(const DeviceArray2D& o) { if (this != &o) { release(); data_ = o.data_; rows_ = o.rows_; cols_ = o.cols_; own_ = false; } return *this; }
    ~DeviceArray2D() { release(); }
    void create(int r, int c) { if (r == rows_ && c == cols_) return; release(); kt::cuda(cudaMalloc((void**)&data_, (size_t)r * c * sizeof(T))); rows_ = r; cols_ = c; own_ = true; }
    void release() { if (own_ && data_) cudaFree(data_); data_ = 0; rows_ = cols_ = 0; own_ = false; }
    void upload(const void* h, size_t host_step, int r, int c) { create(r, c); kt::cuda(cudaMemcpy2D(data_, step(), h, host_step, (size_t)c * sizeof(T), r, cudaMemcpyHostToDevice)); }
    void download(void* h, size_t host_step) const { kt::cuda(cudaMemcpy2D(h, host_step, data_, step(), (size_t)cols_ * sizeof(T), rows_, cudaMemcpyDeviceToHost)); }
    T* ptr(int y = 0) { return data_ + (size_t)y * cols_; } const T* ptr(int y = 0) const { return data_ + (size_t)y * cols_; }
    int cols() const { return cols_; } int rows() const { return rows_; } size_t step() const { return (size_t)cols_ * sizeof(T); }
    bool empty() const { return !data_; }
    operator PtrStep<T>() const { return PtrStep<T>(data_, step()); }
    operator PtrStepSz<T>() const { return PtrStepSz<T>(rows_, cols_, data_, step()); }
private:
    T* data_; int rows_, cols_; bool own_;
};

// ---- POD types of cuda/internal.h ----
struct DataTerm { short2 zero; short2 one; float diff; bool valid; };                       // internal.h:90-96
struct JtJJtrSE3 { float v[27]; float residual, inliers; };                                  // internal.h:98-149 (29 floats)
struct PixelRGB { unsigned char r, g, b; };                                                  // internal.h:151-154
typedef kt_point_xyzrgb PointXYZRGB;                                                         // internal.h:156-184 (32 bytes)
struct Intr { float fx, fy, cx, cy; Intr() : fx(0), fy(0), cx(0), cy(0) {} Intr(float a, float b, float c, float d) : fx(a), fy(b), cx(c), cy(d) {}
    Intr operator()(int level) const { int div = 1 << level; return Intr(fx / div, fy / div, cx / div, cy / div); } };
struct IntrDoublePrecision { double fx, fy, cx, cy; IntrDoublePrecision() : fx(0), fy(0), cx(0), cy(0) {} IntrDoublePrecision(double a, double b, double c, double d) : fx(a), fy(b), cx(c), cy(d) {} };
struct Mat33 { float3 data[3]; };
template <class D, class Matx> D& device_cast(Matx& m) { return *reinterpret_cast<D*>(m.data()); }   // internal.h:481-485

namespace kt { namespace shim {
inline const float* f(const Mat33& m) { return reinterpret_cast<const float*>(&m); }
inline const float* f(const float3& v) { return reinterpret_cast<const float*>(&v); }
inline const float* f(const Intr& k) { return reinterpret_cast<const float*>(&k); }
}}

// ---- operator API: cuda/internal.h:299-536, same names and argument order ----
inline void bilateralFilter(const DeviceArray2D<unsigned short>& src, DeviceArray2D<unsigned short>& dst)
{ dst.create(src.rows(), src.cols()); kt::check(kt_op_bilateral(src.ptr(), dst.ptr(), src.rows(), src.cols(), 0)); }
inline void pyrDown(const DeviceArray2D<unsigned short>& src, DeviceArray2D<unsigned short>& dst)
{ dst.create(src.rows() / 2, src.cols() / 2); kt::check(kt_op_pyrdown(src.ptr(), dst.ptr(), src.rows(), src.cols(), 0)); }
inline void createVMap(const Intr& intr, const DeviceArray2D<unsigned short>& depth, DeviceArray2D<float>& vmap)
{ vmap.create(depth.rows() * 3, depth.cols()); kt::check(kt_op_create_vmap(kt::shim::f(intr), depth.ptr(), vmap.ptr(), depth.rows(), depth.cols(), 0)); }
inline void createNMap(const DeviceArray2D<float>& vmap, DeviceArray2D<float>& nmap)
{ nmap.create(vmap.rows(), vmap.cols()); kt::check(kt_op_create_nmap(vmap.ptr(), nmap.ptr(), vmap.rows() / 3, vmap.cols(), 0)); }
inline void tranformMaps(const DeviceArray2D<float>& vs, const DeviceArray2D<float>& ns, const Mat33& R, const float3& t, DeviceArray2D<float>& vd, DeviceArray2D<float>& nd)
{ vd.create(vs.rows(), vs.cols()); nd.create(vs.rows(), vs.cols()); kt::check(kt_op_transform_maps(vs.ptr(), ns.ptr(), kt::shim::f(R), kt::shim::f(t), vd.ptr(), nd.ptr(), vs.rows() / 3, vs.cols(), 0)); }
inline void resizeVMap(const DeviceArray2D<float>& in, DeviceArray2D<float>& out)
{ out.create((in.rows() / 3 / 2) * 3, in.cols() / 2); kt::check(kt_op_resize_vmap(in.ptr(), out.ptr(), in.rows() / 3, in.cols(), 0)); }
inline void resizeNMap(const DeviceArray2D<float>& in, DeviceArray2D<float>& out)
{ out.create((in.rows() / 3 / 2) * 3, in.cols() / 2); kt::check(kt_op_resize_nmap(in.ptr(), out.ptr(), in.rows() / 3, in.cols(), 0)); }
inline void initVolume(PtrStep<short> volume) { (void)volume; throw kt::Error("initVolume(tsdf) alone: use initVolumes(tsdf, colour) -- both planes are cleared by one kernel"); }
inline void initVolumes(PtrStep<short> tsdf, PtrStep<uchar4> color)
{ kt::check(kt_op_init_volume(tsdf.data, (uint8_t*)color.data, kt::shim::volume_resolution(), 0)); }
// clearVolume{X,Y,Z}[Back] + ...c pairs of the reference collapse into one call per (axis, direction)
inline void clearVolume(int axis, bool back, PtrStep<short> tsdf, PtrStep<uchar4> color, int currentVoxelWrap, int deltaVoxelWrap)
{ kt::check(kt_op_clear_volume(axis, back ? 1 : 0, tsdf.data, (uint8_t*)color.data, kt::shim::volume_resolution(), currentVoxelWrap, deltaVoxelWrap, 0)); }
inline void integrateTsdfVolume(const PtrStepSz<unsigned short>& depth_raw, const Intr& intr, const float3& volume_size, const Mat33& Rcurr_inv, const float3& tcurr,
                                float tranc_dist, PtrStep<short> volume, DeviceArray2D<float>& depthRawScaled, const int3& voxelWrap, PtrStep<uchar4> color_volume,
                                PtrStepSz<uchar3> colors, const DeviceArray2D<float>& nmap_curr, bool angleColor)
{
    depthRawScaled.create(depth_raw.rows, depth_raw.cols);
    kt::check(kt_op_integrate(depth_raw.data, depth_raw.rows, depth_raw.cols, kt::shim::f(intr), kt::shim::f(volume_size), kt::shim::f(Rcurr_inv), kt::shim::f(tcurr), tranc_dist,
                              volume.data, (uint8_t*)color_volume.data, kt::shim::volume_resolution(), &voxelWrap.x, (const uint8_t*)colors.data, nmap_curr.ptr(), angleColor ? 1 : 0,
                              depthRawScaled.ptr(), 0));
}
inline void raycast(const Intr& intr, const Mat33& Rcurr, const float3& tcurr, float tranc_dist, const float3& volume_size, const PtrStep<short>& volume,
                    DeviceArray2D<float>& vmap, DeviceArray2D<float>& nmap, const int3& voxelWrap, DeviceArray2D<uchar4>& vmap_curr_color, PtrStep<uchar4> color_volume)
{
    kt::check(kt_op_raycast(kt::shim::f(intr), kt::shim::f(Rcurr), kt::shim::f(tcurr), tranc_dist, kt::shim::f(volume_size), volume.data, kt::shim::volume_resolution(),
                            vmap.ptr(), nmap.ptr(), vmap.rows() / 3, vmap.cols(), &voxelWrap.x, (uint8_t*)vmap_curr_color.ptr(), (const uint8_t*)color_volume.data, 0));
}
inline size_t extractCloudSlice(const PtrStep<short>& volume, const float3& volume_size, PtrSz<PointXYZRGB> output, int3 voxelWrap, PtrStep<uchar4>& color_volume,
                                int minX, int maxX, int minY, int maxY, int minZ, int maxZ, int subsample, int3 realVoxelWrap)
{
    size_t n = 0;
    kt::check(kt_op_extract_slice(volume.data, kt::shim::f(volume_size), kt::shim::volume_resolution(), output.data, output.size, &voxelWrap.x, (const uint8_t*)color_volume.data,
                                  minX, maxX, minY, maxY, minZ, maxZ, subsample, &realVoxelWrap.x, &n, 0));
    return n;
}
inline void icpStep(const Mat33& Rcurr, const float3& tcurr, const DeviceArray2D<float>& vmap_curr, const DeviceArray2D<float>& nmap_curr, const Mat33& Rprev_inv, const float3& tprev,
                    const Intr& intr, const DeviceArray2D<float>& vmap_g_prev, const DeviceArray2D<float>& nmap_g_prev, float distThres, float angleThres,
                    DeviceArray<JtJJtrSE3>& /*sum*/, DeviceArray<JtJJtrSE3>& /*out*/, float* matrixA_host, float* vectorB_host, float* residual_host, int /*threads*/, int /*blocks*/)
{
    kt::check(kt_op_icp_step(kt::shim::f(Rcurr), kt::shim::f(tcurr), vmap_curr.ptr(), nmap_curr.ptr(), kt::shim::f(Rprev_inv), kt::shim::f(tprev), kt::shim::f(intr),
                             vmap_g_prev.ptr(), nmap_g_prev.ptr(), vmap_curr.rows() / 3, vmap_curr.cols(), distThres, angleThres, matrixA_host, vectorB_host, residual_host, 0));
}
inline void shortDepthToMetres(const DeviceArray2D<unsigned short>& src, DeviceArray2D<float>& dst, int cutOff)
{ kt::check(kt_op_short_depth_to_metres(src.ptr(), dst.ptr(), dst.rows(), dst.cols(), cutOff, 0)); }
inline void pyrDownGaussF(const DeviceArray2D<float>& src, DeviceArray2D<float>& dst)
{ dst.create(src.rows() / 2, src.cols() / 2); kt::check(kt_op_pyrdown_gauss_f(src.ptr(), dst.ptr(), src.rows(), src.cols(), 0)); }
inline void imageBGRToIntensity(const DeviceArray2D<PixelRGB>& src, DeviceArray2D<unsigned char>& dst)
{ kt::check(kt_op_bgr_to_intensity((const uint8_t*)src.ptr(), dst.ptr(), dst.rows(), dst.cols(), 0)); }
inline void pyrDownUcharGauss(const DeviceArray2D<unsigned char>& src, DeviceArray2D<unsigned char>& dst)
{ dst.create(src.rows() / 2, src.cols() / 2); kt::check(kt_op_pyrdown_uchar_gauss(src.ptr(), dst.ptr(), src.rows(), src.cols(), 0)); }
inline void computeDerivativeImages(DeviceArray2D<unsigned char>& src, DeviceArray2D<short>& dx, DeviceArray2D<short>& dy)
{ kt::check(kt_op_derivative_images(src.ptr(), dx.ptr(), dy.ptr(), src.rows(), src.cols(), 0)); }
inline void projectToPointCloud(const DeviceArray2D<float>& depth, const DeviceArray2D<float3>& cloud, IntrDoublePrecision& intrinsics, const int& level)
{ const double k[4] = {intrinsics.fx, intrinsics.fy, intrinsics.cx, intrinsics.cy}; kt::check(kt_op_project_to_point_cloud(depth.ptr(), (float*)cloud.ptr(), depth.rows(), depth.cols(), k, level, 0)); }
inline void computeRgbResidual(const float& minScale, const DeviceArray2D<short>& dIdx, const DeviceArray2D<short>& dIdy, const DeviceArray2D<float>& lastDepth,
                               const DeviceArray2D<float>& nextDepth, const DeviceArray2D<unsigned char>& lastImage, const DeviceArray2D<unsigned char>& nextImage,
                               DeviceArray2D<DataTerm>& corresImg, DeviceArray<int2>& /*sumResidual*/, const float maxDepthDelta, const float3& kt_, const Mat33& krkinv,
                               int& sigmaSum, int& count, int /*threads*/, int /*blocks*/)
{
    kt::check(kt_op_rgb_residual(minScale, dIdx.ptr(), dIdy.ptr(), lastDepth.ptr(), nextDepth.ptr(), lastImage.ptr(), nextImage.ptr(), corresImg.ptr(), nextImage.rows(), nextImage.cols(),
                                 maxDepthDelta, kt::shim::f(kt_), kt::shim::f(krkinv), &sigmaSum, &count, 0));
}
inline void rgbStep(const DeviceArray2D<DataTerm>& corresImg, const float& sigma, const DeviceArray2D<float3>& cloud, const float& fx, const float& fy, const DeviceArray2D<short>& dIdx,
                    const DeviceArray2D<short>& dIdy, const float& sobelScale, DeviceArray<JtJJtrSE3>& /*sum*/, DeviceArray<JtJJtrSE3>& /*out*/, float* matrixA_host, float* vectorB_host,
                    int /*threads*/, int /*blocks*/)
{
    kt::check(kt_op_rgb_step(corresImg.ptr(), sigma, (const float*)cloud.ptr(), fx, fy, dIdx.ptr(), dIdy.ptr(), sobelScale, corresImg.rows(), corresImg.cols(), matrixA_host, vectorB_host, 0));
}

// ---- TsdfVolume / ColorVolume (TSDFVolume.h:68-158, TSDFVolume.cpp:60-215, ColorVolume.h:65-94) without Eigen ----
// The reference constructs TsdfVolume(Eigen::Vector3i) + ColorVolume(tsdf) and hands `data()` of both to integrateTsdfVolume / raycast /
// extractCloudSlice / clearVolume*; these two classes own the same two device planes (short[V^3], uchar4[V^3] with the weight in .w) and
// keep the reference's parameter logic (truncation distance clamp, voxel size).
class ColorVolume;
class TsdfVolume {
public:
    enum { DEFAULT_CLOUD_BUFFER_SIZE = 10 * 1000 * 1000 };
    // resolution: voxels per side (the reference asserts a cube); volumeSize: Volume::get().getVolumeSize() in metres
    explicit TsdfVolume(int resolution, float volumeSize = 6.f) : res_(resolution), tranc_dist_(0.03f)
    {
        kt::shim::set_volume_resolution(resolution);
        volume_.create(resolution * resolution, resolution);
        size_ = make_float3(volumeSize, volumeSize, volumeSize);
        setTsdfTruncDist(0.03f);                                     // default_tranc_dist (TSDFVolume.cpp:73)
        kt::cuda(cudaMemset(volume_.ptr(), 0, (size_t)resolution * resolution * resolution * sizeof(short)));
    }
    void setSize(const float3& size) { size_ = size; setTsdfTruncDist(tranc_dist_); }
    void setTsdfTruncDist(float distance)                          // TSDFVolume.cpp:89-97: at least 2.1 voxels
    {
        const float cx = size_.x / res_, cy = size_.y / res_, cz = size_.z / res_;
        const float m = cx > cy ? (cx > cz ? cx : cz) : (cy > cz ? cy : cz);
        tranc_dist_ = distance > 2.1f * m ? distance : 2.1f * m;
    }
    DeviceArray2D<short> data() const { return volume_; }
    const float3& getSize() const { return size_; }
    int getResolution() const { return res_; }
    float3 getVoxelSize() const { return make_float3(size_.x / res_, size_.y / res_, size_.z / res_); }
    float getTsdfTruncDist() const { return tranc_dist_; }
    // reset(): the reference clears the TSDF plane here and the colour plane in ColorVolume::reset(); one kernel clears both
    inline void reset(ColorVolume& color);
    // fetchCloud (TSDFVolume.cpp:137-172): returns a non-owning view of the filled part of cloud_buffer
    DeviceArray<PointXYZRGB> fetchCloud(DeviceArray<PointXYZRGB>& cloud_buffer, int3& voxelWrap, PtrStep<uchar4> color_volume, int minX, int maxX,
                                        int minY, int maxY, int minZ, int maxZ, int3 realVoxelWrap, int subsample = 1) const
    {
        if (cloud_buffer.empty()) cloud_buffer.create(DEFAULT_CLOUD_BUFFER_SIZE);
        const size_t n = extractCloudSlice(PtrStep<short>(const_cast<short*>(volume_.ptr()), volume_.step()), size_,
                                           PtrSz<PointXYZRGB>(cloud_buffer.ptr(), cloud_buffer.size()), voxelWrap, color_volume,
                                           minX, maxX, minY, maxY, minZ, maxZ, subsample, realVoxelWrap);
        return DeviceArray<PointXYZRGB>(cloud_buffer.ptr(), n);
    }
    // downloadTsdf (TSDFVolume.cpp:174-185): value / DIVISOR per voxel, index res*res*z + res*y + x in STORAGE order.  The reference still
    // reads the volume as the short2 {tsdf, weight} pairs of its PCL ancestor (it copies cols*sizeof(int) bytes per row of a short
    // volume); here the volume is read as what it is, one short per voxel.
    void downloadTsdf(std::vector<float>& tsdf) const
    {
        std::vector<short> raw((size_t)res_ * res_ * res_);
        kt::cuda(cudaMemcpy(&raw[0], volume_.ptr(), raw.size() * sizeof(short), cudaMemcpyDeviceToHost));
        tsdf.resize(raw.size());
        for (size_t i = 0; i < raw.size(); ++i) tsdf[i] = (float)raw[i] / 32767.f;      // DIVISOR (cuda/device.hpp)
    }
    // downloadTsdfAndWeighs (TSDFVolume.cpp:187-201): the weights live in the .w byte of the colour volume
    inline void downloadTsdfAndWeighs(const ColorVolume& color, std::vector<float>& tsdf, std::vector<short>& weights) const;
    // saveTsdfToDisk (TSDFVolume.cpp:203-215+): <name>_tsdf.bin (float) and <name>_weights.bin (short)
    inline void saveTsdfToDisk(const ColorVolume& color, const std::string& filename) const;
private:
    TsdfVolume(const TsdfVolume&); TsdfVolume& operator=(const TsdfVolume&);
    int res_; float3 size_; DeviceArray2D<short> volume_; float tranc_dist_;
};

class ColorVolume {
public:
    explicit ColorVolume(const TsdfVolume& tsdf) : res_(tsdf.getResolution())
    {
        color_volume_.create(res_ * res_, res_);
        reset();
    }
    void reset() { kt::cuda(cudaMemset(color_volume_.ptr(), 0, (size_t)res_ * res_ * res_ * sizeof(int))); }
    DeviceArray2D<int> data() const { return color_volume_; }         // declared int, used as uchar4 (ColorVolume.h:93, Q11)
    PtrStep<uchar4> view() const { return PtrStep<uchar4>((uchar4*)const_cast<int*>(color_volume_.ptr()), color_volume_.step()); }
private:
    ColorVolume(const ColorVolume&); ColorVolume& operator=(const ColorVolume&);
    int res_; DeviceArray2D<int> color_volume_;
};

inline void TsdfVolume::reset(ColorVolume& color)
{ initVolumes(PtrStep<short>(volume_.ptr(), volume_.step()), color.view()); }
inline void TsdfVolume::downloadTsdfAndWeighs(const ColorVolume& color, std::vector<float>& tsdf, std::vector<short>& weights) const
{
    downloadTsdf(tsdf);
    std::vector<int> raw((size_t)res_ * res_ * res_);
    kt::cuda(cudaMemcpy(&raw[0], color.data().ptr(), raw.size() * sizeof(int), cudaMemcpyDeviceToHost));
    weights.resize(raw.size());
    for (size_t i = 0; i < raw.size(); ++i) weights[i] = (short)(((unsigned int)raw[i]) >> 24);      // uchar4 .w
}
inline void TsdfVolume::saveTsdfToDisk(const ColorVolume& color, const std::string& filename) const
{
    std::vector<float> tsdf; std::vector<short> weights;
    downloadTsdfAndWeighs(color, tsdf, weights);
    FILE* fp = std::fopen((filename + "_tsdf.bin").c_str(), "wb");
    if (!fp) throw kt::Error("saveTsdfToDisk: cannot open " + filename + "_tsdf.bin");
    std::fwrite(&tsdf[0], sizeof(float), tsdf.size(), fp); std::fclose(fp);
    fp = std::fopen((filename + "_weights.bin").c_str(), "wb");
    if (!fp) throw kt::Error("saveTsdfToDisk: cannot open " + filename + "_weights.bin");
    std::fwrite(&weights[0], sizeof(short), weights.size(), fp); std::fclose(fp);
}

// ---- CloudSlice (CloudSlice.h:28-129): the record handed to the backend on every volume shift, without the PCL / Eigen types ----
struct CloudSliceB200 {
    enum Dimension { XPlus, XMinus, YPlus, YMinus, ZPlus, ZMinus, FIRST, FINAL, TSDF };      // CloudSlice.h:33-36
    enum Odometry { ICP, GROUNDTRUTH, RGBD, FAIL };                                          // CloudSlice.h:38-41
    std::vector<PointXYZRGB> cloud;            // pcl::PointCloud<pcl::PointXYZRGB>::points, same 32-byte layout
    Dimension dimension; Odometry odometry;
    float cameraTranslation[3]; float cameraRotation[9];        // Eigen::Vector3f / Matrix<float,3,3,RowMajor>
    uint64_t utime;
};

// ---- KintinuousTracker facade (KintinuousTracker.h:85-172) without Eigen / cv / pcl / boost types ----
class KintinuousTrackerB200 {
public:
    // K = {fx, fy, cx, cy} of the cv::Mat the reference constructor takes (KintinuousTracker.cpp:86-89)
    explicit KintinuousTrackerB200(const kt_config& cfg) : ctx_(0) { kt::check(kt_create(&cfg, &ctx_)); }
    ~KintinuousTrackerB200() { kt_destroy(ctx_); }
    // processFrame(depth, colors, rgbHost, depthHost, utime, ...): host buffers, the upload is done inside
    kt_pose processFrame(const unsigned short* depthHost, const unsigned char* rgbHost, uint64_t utime) { kt_pose p; kt::check(kt_process_frame(ctx_, depthHost, rgbHost, utime, &p)); return p; }
    kt_pose processFrame(const DeviceArray2D<unsigned short>& depth, const DeviceArray2D<PixelRGB>& colors, uint64_t utime)
    { kt_pose p; kt::check(kt_process_frame_device(ctx_, depth.ptr(), (const uint8_t*)colors.ptr(), utime, &p)); return p; }
    // optional hint (no counterpart in the reference): start the next frame's upload and pose-independent front end now
    void prefetchFrame(const unsigned short* depth, const unsigned char* rgb) { kt::check(kt_prefetch_frame(ctx_, depth, rgb)); }
    void finalise() { kt::check(kt_finalise(ctx_)); }
    void reset() { kt::check(kt_reset(ctx_)); }
    float getVoxelSize() const { return kt_get_voxel_size(ctx_); }
    void setOverlap(int o) { kt::check(kt_set_overlap(ctx_, o)); }
    void setParked(bool p) { kt::check(kt_set_parked(ctx_, p ? 1 : 0)); }
    kt_pose getLastPose() const { kt_pose p; kt::check(kt_get_pose(ctx_, &p)); return p; }   // getLastRotation / getLastTranslation / getVolumeOffset
    int numCloudSlices() const { return kt_num_slices(ctx_); }                                  // getCloudSlices().size()
    std::vector<PointXYZRGB> getCloudSlice(int i, int* dimension = 0) const
    {
        size_t n = 0; kt::check(kt_get_slice(ctx_, i, 0, 0, &n, dimension, 0));
        std::vector<PointXYZRGB> v(n);
        if (n) kt::check(kt_get_slice(ctx_, i, &v[0], n, &n, dimension, 0));
        return v;
    }
    // the whole CloudSlice record of slice i (cloud, dimension, odometry kind, camera pose at hand-over, timestamp)
    CloudSliceB200 getCloudSliceRecord(int i) const
    {
        kt_slice_info info; kt::check(kt_get_slice_info(ctx_, i, &info));
        CloudSliceB200 s;
        s.cloud.resize(info.count);
        size_t n = info.count;
        if (n) kt::check(kt_get_slice(ctx_, i, &s.cloud[0], n, &n, 0, 0));
        s.dimension = (CloudSliceB200::Dimension)info.dimension; s.odometry = (CloudSliceB200::Odometry)info.odometry;
        for (int k = 0; k < 3; ++k) s.cameraTranslation[k] = info.camera_t[k];
        for (int k = 0; k < 9; ++k) s.cameraRotation[k] = info.camera_R[k];
        s.utime = info.utime;
        return s;
    }
    kt_ctx* handle() { return ctx_; }
private:
    KintinuousTrackerB200(const KintinuousTrackerB200&); KintinuousTrackerB200& operator=(const KintinuousTrackerB200&);
    kt_ctx* ctx_;
};

#endif // KINTINUOUS_B200_SHIM_HPP_
