// Host build of the product's on-device Gauss-Newton step (kintinuous_b200/csrc/kt_solve.cuh) for CPU-side unit tests:
// the SAME source that runs inside icp_frame_kernel / rgbd_frame_kernel, compiled for the host and exported with a C ABI.
//   nvcc -std=c++17 -shared -Xcompiler -fPIC -I kintinuous_b200/csrc -o tests/cpp/_build/libkt_solve_host.so tests/cpp/solve_host.cu
#include "kt_solve.cuh"

extern "C" {
void kts_ldlt6_solve(const double* A, const double* b, double* x) { kt::ldlt6_solve(A, b, x); }
void kts_rodrigues(const double* r, double* R) { kt::rodrigues(r, R); }
void kts_mat3f_inverse(const float* m, float* r) { kt::mat3f_inverse(m, r); }
void kts_unpack(const float* sums27, float* A36, float* b6) { kt::unpack_normal_equations(sums27, A36, b6); }
// one Gauss-Newton update: resultRt (4x4 double, in/out), previous pose -> current pose
void kts_update(const double* A, const double* b, double* resultRt, const float* Rprev, const float* tprev, float* Rcurr, float* tcurr)
{ kt::gauss_newton_update_p(A, b, resultRt, Rprev, tprev, Rcurr, tcurr); }
// the latency-trimmed forms the whole-frame kernels use (series Rodrigues, 3-row product; the reciprocal seed is device-only)
void kts_ldlt6_solve_fast(const double* A, const double* b, double* x) { kt::ldlt6_solve_fast(A, b, x); }
void kts_rodrigues_fast(const double* r, double* R) { kt::rodrigues_fast(r, R); }
void kts_update_fast(const double* A, const double* b, double* resultRt, const float* Rprev, const float* tprev, float* Rcurr, float* tcurr)
{ kt::gauss_newton_update_fast(A, b, resultRt, Rprev, tprev, Rcurr, tcurr); }
}
