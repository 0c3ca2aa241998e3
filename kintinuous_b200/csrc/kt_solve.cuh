// kintinuous_b200 -- on-device Gauss-Newton step: 6x6 solve, se(3) exponential and pose update.
//
// Replaces the HOST side of every odometry iteration in the reference, so that the coarse-to-fine
// loop runs without a device->host round trip (the reference does sync + 116-byte D2H + host LDLT
// 19 times per frame):
//   unpack 27 sums -> A (6x6, symmetric), b (6)          cuda/reduce.cu:404-415
//   x = A.ldlt().solve(b) in double                       ICPOdometry.cpp:127-131, RGBDOdometry.cpp:316-326
//   currRt = [Rodrigues(x[3..5]) | x[0..2]]               OdometryProvider.h:54-68 (cv::Rodrigues, 64F)
//   resultRt = currRt * resultRt                          ICPOdometry.cpp:144
//   [Rcurr|tcurr] = [Rprev|tprev] * inverse([rot|trans])  ICPOdometry.cpp:146-178 (Eigen::Isometry3f, float)
// One thread executes this (a few hundred dependent FP64 ops, ~1-2 us); it is called by the last
// CTA of the reduction kernel.
#pragma once
#include "kt_ops.h"
#include <float.h>

namespace kt {

// sums: 27 upper-triangular products in the order aa ab ac ad ae af ag bb ... fg (cuda/internal.h:101-106)
__device__ __forceinline__ void unpack_normal_equations(const float* sums, float* A, float* b)
{
    int shift = 0;
    for (int i = 0; i < 6; ++i)
        for (int j = i; j < 7; ++j) {
            float value = sums[shift++];
            if (j == 6) b[i] = value;
            else A[j * 6 + i] = A[i * 6 + j] = value;
        }
}

// LDL^T with diagonal pivoting (largest |d| first), as Eigen::LDLT does; double precision.
__device__ inline void ldlt6_solve(const double* Ain, const double* b, double* x)
{
    const int n = 6;
    double a[36];
    int perm[6];
    for (int i = 0; i < 36; ++i) a[i] = Ain[i];
    for (int i = 0; i < n; ++i) perm[i] = i;
    for (int k = 0; k < n; ++k) {
        int piv = k; double best = fabs(a[k * n + k]);
        for (int i = k + 1; i < n; ++i) { double v = fabs(a[i * n + i]); if (v > best) { best = v; piv = i; } }
        if (piv != k) {
            for (int j = 0; j < n; ++j) { double t = a[k * n + j]; a[k * n + j] = a[piv * n + j]; a[piv * n + j] = t; }
            for (int i = 0; i < n; ++i) { double t = a[i * n + k]; a[i * n + k] = a[i * n + piv]; a[i * n + piv] = t; }
            int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
        }
        double d = a[k * n + k];
        if (d == 0.0) continue;
        for (int i = k + 1; i < n; ++i) a[i * n + k] /= d;
        for (int i = k + 1; i < n; ++i)
            for (int j = k + 1; j <= i; ++j) {
                a[i * n + j] -= a[i * n + k] * d * a[j * n + k];
                a[j * n + i] = a[i * n + j];
            }
    }
    double y[6];
    for (int i = 0; i < n; ++i) y[i] = b[perm[i]];
    for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j) y[i] -= a[i * n + j] * y[j];
    for (int i = 0; i < n; ++i) { double d = a[i * n + i]; y[i] = (fabs(d) > DBL_MIN) ? y[i] / d : 0.0; }
    for (int i = n - 1; i >= 0; --i) for (int j = i + 1; j < n; ++j) y[i] -= a[j * n + i] * y[j];
    for (int i = 0; i < n; ++i) x[perm[i]] = y[i];
}

// cv::Rodrigues, rotation vector -> matrix, double (OpenCV 2.4.9 semantics)
__device__ inline void rodrigues(const double* r, double* R)
{
    double rx = r[0], ry = r[1], rz = r[2];
    double theta = sqrt(rx * rx + ry * ry + rz * rz);
    if (theta < DBL_EPSILON) {
        for (int k = 0; k < 9; ++k) R[k] = (k % 4 == 0) ? 1.0 : 0.0;
        return;
    }
    double c = cos(theta), s = sin(theta), c1 = 1.0 - c;
    double itheta = 1.0 / theta;
    rx *= itheta; ry *= itheta; rz *= itheta;
    const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
    double rx_[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
    for (int k = 0; k < 9; ++k) R[k] = c * I[k] + c1 * rrt[k] + s * rx_[k];
}

__device__ __forceinline__ void mat3f_inverse(const float* m, float* r)   // Eigen 3x3 inverse (cofactors / det)
{
#define KT_M(i, j) m[(i) * 3 + (j)]
#define KT_COF(i, j) (KT_M(((i) + 1) % 3, ((j) + 1) % 3) * KT_M(((i) + 2) % 3, ((j) + 2) % 3) - KT_M(((i) + 1) % 3, ((j) + 2) % 3) * KT_M(((i) + 2) % 3, ((j) + 1) % 3))
    float c00 = __fsub_rn(__fmul_rn(KT_M(1, 1), KT_M(2, 2)), __fmul_rn(KT_M(1, 2), KT_M(2, 1)));
    float c10 = __fsub_rn(__fmul_rn(KT_M(2, 1), KT_M(0, 2)), __fmul_rn(KT_M(2, 2), KT_M(0, 1)));
    float c20 = __fsub_rn(__fmul_rn(KT_M(0, 1), KT_M(1, 2)), __fmul_rn(KT_M(0, 2), KT_M(1, 1)));
    float det = __fadd_rn(__fadd_rn(__fmul_rn(c00, KT_M(0, 0)), __fmul_rn(c10, KT_M(1, 0))), __fmul_rn(c20, KT_M(2, 0)));
    float invdet = __fdiv_rn(1.0f, det);
    float c01 = __fsub_rn(__fmul_rn(KT_M(1, 2), KT_M(2, 0)), __fmul_rn(KT_M(1, 0), KT_M(2, 2)));
    float c11 = __fsub_rn(__fmul_rn(KT_M(2, 2), KT_M(0, 0)), __fmul_rn(KT_M(2, 0), KT_M(0, 2)));
    float c21 = __fsub_rn(__fmul_rn(KT_M(0, 2), KT_M(1, 0)), __fmul_rn(KT_M(0, 0), KT_M(1, 2)));
    float c02 = __fsub_rn(__fmul_rn(KT_M(1, 0), KT_M(2, 1)), __fmul_rn(KT_M(1, 1), KT_M(2, 0)));
    float c12 = __fsub_rn(__fmul_rn(KT_M(2, 0), KT_M(0, 1)), __fmul_rn(KT_M(2, 1), KT_M(0, 0)));
    float c22 = __fsub_rn(__fmul_rn(KT_M(0, 0), KT_M(1, 1)), __fmul_rn(KT_M(0, 1), KT_M(1, 0)));
    r[0] = __fmul_rn(c00, invdet); r[1] = __fmul_rn(c10, invdet); r[2] = __fmul_rn(c20, invdet);
    r[3] = __fmul_rn(c01, invdet); r[4] = __fmul_rn(c11, invdet); r[5] = __fmul_rn(c21, invdet);
    r[6] = __fmul_rn(c02, invdet); r[7] = __fmul_rn(c12, invdet); r[8] = __fmul_rn(c22, invdet);
#undef KT_COF
#undef KT_M
}

// IEEE (non-contracted) float helpers: the reference does this part on the HOST CPU, without FMA.
__device__ __forceinline__ float dot3_rn(float a0, float a1, float a2, float b0, float b1, float b2)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1)), __fmul_rn(a2, b2));
}

// Solve for the increment and update st->resultRt / Rcurr / tcurr.  A, b in double.
__device__ inline void gauss_newton_update(const double* dA, const double* db, OdomState* st)
{
    double x[6];
    ldlt6_solve(dA, db, x);
    double R[9];
    rodrigues(x + 3, R);
    double cur[16] = {R[0], R[1], R[2], x[0], R[3], R[4], R[5], x[1], R[6], R[7], R[8], x[2], 0, 0, 0, 1};
    double res[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += cur[i * 4 + k] * st->resultRt[k * 4 + j];
            res[i * 4 + j] = s;
        }
    for (int k = 0; k < 16; ++k) st->resultRt[k] = res[k];
    // float part (Eigen::Isometry3f): inverse of [rot|tr] is [rot^T | -rot^T tr]; then Rprev * that.
    float rot[9], tr[3];
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) rot[i * 3 + j] = (float)res[i * 4 + j]; tr[i] = (float)res[i * 4 + 3]; }
    float tinv[3];
    for (int i = 0; i < 3; ++i) tinv[i] = -dot3_rn(rot[0 * 3 + i], rot[1 * 3 + i], rot[2 * 3 + i], tr[0], tr[1], tr[2]);
    const float* Rp = st->Rprev;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j)   // (Rprev * rot^T)(i,j) = sum_k Rprev(i,k) * rot(j,k)
            st->Rcurr[i * 3 + j] = dot3_rn(Rp[i * 3 + 0], Rp[i * 3 + 1], Rp[i * 3 + 2], rot[j * 3 + 0], rot[j * 3 + 1], rot[j * 3 + 2]);
        st->tcurr[i] = __fadd_rn(dot3_rn(Rp[i * 3 + 0], Rp[i * 3 + 1], Rp[i * 3 + 2], tinv[0], tinv[1], tinv[2]), st->tprev[i]);
    }
}

} // namespace kt
