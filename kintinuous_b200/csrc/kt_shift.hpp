// kintinuous_b200 -- host-side bookkeeping of the shifting volume, free of CUDA so that it also builds into a CPU unit test
// (tests/cpp/shift_host.cpp, tests/test_shift_logic.py).  Each function restates a few lines of KintinuousTracker.cpp / TSDFVolume.cpp.
#pragma once
#include <algorithm>
#include <climits>
#include <cmath>

namespace kt {

// TsdfVolume::setTsdfTruncDist with the tracker's request (KintinuousTracker.cpp:112 asks for max(0.01, size / 100); TSDFVolume.cpp:96
// keeps it at or above 2.1 voxels)
inline float trunc_dist_for(float size, float voxel)
{
    const float def = std::max(0.01f, size / 100.f);
    return std::max(def, 2.1f * voxel);
}

// KintinuousTracker::vWrapCopyUpdate (.cpp:1075-1085): the non-negative alias of the signed, unbounded voxel wrap
inline void vwrap_nonneg(const int* voxelWrap, int V, int* w)
{
    for (int i = 0; i < 3; ++i) { w[i] = voxelWrap[i]; if (w[i] < 0) w[i] = V - ((-w[i]) % V); }
}

// currentGlobalCamera (.cpp:581-596): camera position in the world the volume travels through
inline float global_camera(float basis, float size, int voxelWrap, float voxel, float t)
{
    float g = basis - size * 0.5f;
    g += voxelWrap * voxel;
    g += t - basis;
    return g;
}

// Whole voxels the camera has moved away from the volume centre, clamped to +-thresh (.cpp:636-667); thresh = INT_MAX when parked
inline void shift_steps(const float* currentTranslation, float voxel, int thresh, int* trans)
{
    for (int i = 0; i < 3; ++i) {
        const int f = (int)std::floor(currentTranslation[i] / voxel);
        trans[i] = (f < 0) ? std::max(-thresh, f) : std::min(thresh, f);
    }
}

// Box [lo, hi) of the slab that leaves the volume when it moves by n voxels along `axis` (x .cpp:675-723, y :729-777, z :783-831).
// Returns +1: the front slab leaves (clearVolume*), -1: the back slab leaves (clearVolume*Back), 0: no shift on this axis.
// The ZMinus slab sits one plane lower than its X / Y equivalents (.cpp:805, Q12).
inline int shift_box(int axis, int n, int thresh, int overlap, int V, int* lo, int* hi)
{
    for (int i = 0; i < 3; ++i) { lo[i] = 0; hi[i] = V; }
    if (n >= thresh) { lo[axis] = 0; hi[axis] = n + 1 + overlap; return 1; }
    if (n <= -thresh) {
        if (axis < 2) { lo[axis] = V + (n - overlap); hi[axis] = V; }
        else { lo[axis] = V + (n - overlap) - 1; hi[axis] = V - 1; }
        return -1;
    }
    return 0;
}

// CloudSlice::Dimension of a shift by vt voxels (CloudSlice.h:33-36; mutexOutCloudBuffer, .cpp:1156-1208)
inline int slice_dimension(const int* vt)
{
    return vt[0] > 0 ? 0 : vt[0] < 0 ? 1 : vt[1] > 0 ? 2 : vt[1] < 0 ? 3 : vt[2] > 0 ? 4 : 5;
}

} // namespace kt
