// Host build of kintinuous_b200/csrc/kt_posegraph.hpp for tests/test_posegraph.py.
#include "kt_posegraph.hpp"
extern "C" {
void kth_quaternion(const float* R9, float* q4) { kt::quaternion_from_rotation(R9, q4); }
int kth_pose_line(unsigned long long ts, const float* t3, const float* R9, char* buf, int cap) { return kt::format_pose_line(ts, t3, R9, buf, (size_t)cap); }
}
