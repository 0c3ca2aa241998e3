// Host build of the product's volume-shift bookkeeping (kintinuous_b200/csrc/kt_shift.hpp) for tests/test_shift_logic.py.
//   g++ -std=c++14 -O1 -shared -fPIC -I kintinuous_b200/csrc -o tests/cpp/_build/libkt_shift_host.so tests/cpp/shift_host.cpp
#include "kt_shift.hpp"

extern "C" {
float kth_trunc_dist(float size, float voxel) { return kt::trunc_dist_for(size, voxel); }
void kth_vwrap_nonneg(const int* wrap, int V, int* out) { kt::vwrap_nonneg(wrap, V, out); }
float kth_global_camera(float basis, float size, int wrap, float voxel, float t) { return kt::global_camera(basis, size, wrap, voxel, t); }
void kth_shift_steps(const float* ct, float voxel, int thresh, int* trans) { kt::shift_steps(ct, voxel, thresh, trans); }
int kth_shift_box(int axis, int n, int thresh, int overlap, int V, int* lo, int* hi) { return kt::shift_box(axis, n, thresh, overlap, V, lo, hi); }
int kth_slice_dimension(const int* vt) { return kt::slice_dimension(vt); }
}
