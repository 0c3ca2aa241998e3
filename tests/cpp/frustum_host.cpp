// Host build of kintinuous_b200/csrc/kt_frustum.hpp for tests/test_frustum_box.py.
//   g++ -std=c++14 -O1 -shared -fPIC -I kintinuous_b200/csrc -o tests/cpp/_build/libkt_frustum_host.so tests/cpp/frustum_host.cpp
#include "kt_frustum.hpp"

extern "C" {
int kth_frustum_box(const float* Rinv, const float* t, const float* k4, int rows, int cols, int V, const float* cell, int* lo3, int* hi3)
{
    const kt::VoxelBox b = kt::frustum_voxel_box(Rinv, t, k4, rows, cols, V, cell);
    for (int i = 0; i < 3; ++i) { lo3[i] = b.lo[i]; hi3[i] = b.hi[i]; }
    return b.empty ? 1 : 0;
}
void kth_cyclic_tile_range(int lo, int hi, int wrap, int V, int tile, int* first, int* n) { kt::cyclic_tile_range(lo, hi, wrap, V, tile, first, n); }
}
