// Host build of kt_replay.cuh for the CPU unit test (tests/test_replay_add_host.py).
//   g++ -O1 -ffp-contract=off -shared -fPIC -I kintinuous_b200/csrc -o tests/cpp/_build/libkt_replay_host.so tests/cpp/replay_host.cpp
#define KT_HD
#define __forceinline__ inline
#include "kt_replay.cuh"
extern "C" {
float ktr_replay_add(float x, float a, int k) { return kt::replay_add(x, a, k); }
float ktr_loop_add(float x, float a, int k) { volatile float v = x; for (int i = 0; i < k; ++i) v = v + a; return v; }
float ktr_replay_fma(float x, float m, float f, int k) { return kt::replay_fma(x, m, f, k); }
float ktr_loop_fma(float x, float m, float f, int k) { volatile float v = x; for (int i = 0; i < k; ++i) v = fmaf(m, f, v); return v; }
int ktr_check_many_fma(const float* x, const float* m, const float* f, const int* k, int n)
{
    int bad = 0;
    for (int i = 0; i < n; ++i) {
        float r0 = ktr_loop_fma(x[i], m[i], f[i], k[i]), r1 = kt::replay_fma(x[i], m[i], f[i], k[i]);
        unsigned u0, u1; memcpy(&u0, &r0, 4); memcpy(&u1, &r1, 4);
        if (u0 != u1) ++bad;
    }
    return bad;
}
// many cases at once: returns the number of mismatching results (bitwise)
int ktr_check_many(const float* x, const float* a, const int* k, int n)
{
    int bad = 0;
    for (int i = 0; i < n; ++i) {
        float r0 = ktr_loop_add(x[i], a[i], k[i]), r1 = kt::replay_add(x[i], a[i], k[i]);
        unsigned u0, u1; memcpy(&u0, &r0, 4); memcpy(&u1, &r1, 4);
        if (u0 != u1) ++bad;
    }
    return bad;
}
}
