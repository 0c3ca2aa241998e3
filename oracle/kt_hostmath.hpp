// TEST INFRASTRUCTURE ONLY (oracle/): never linked into, imported by or executed from the
// product path (kintinuous_b200/). Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use anything under oracle/.
//
// Host-side small linear algebra restating the third-party arithmetic that the reference's
// host classes take from Eigen / OpenCV (neither is installed here):
//   * Eigen::Matrix3f::inverse()            (ICPOdometry.cpp:81, KintinuousTracker.cpp:627)
//   * Eigen::LDLT<Matrix<double,6,6>>       (ICPOdometry.cpp:131, RGBDOdometry.cpp:320,325)
//   * cv::Rodrigues (vector -> matrix, 64F) (OdometryProvider.h:54-68; OpenCV 2.4.9, build.sh:57)
//   * cv::Mat 4x4 64F product               (ICPOdometry.cpp:144)
//   * Eigen::Isometry3f compose / inverse   (ICPOdometry.cpp:164-178)
//   * cv::Mat::inv(DECOMP_SVD) of a rigid 4x4, K*R*K^-1 (RGBDOdometry.cpp:209-231)
// Parity note: these are restatements of published algorithms, compiled with
// -ffp-contract=off; they agree with Eigen/OpenCV to rounding (~1e-7 float, ~1e-15 double).
#pragma once
#include <cmath>
#include <cfloat>
#include <cstring>

namespace kto {

struct Mat3f { float m[9]; };   // row-major, like Eigen::Matrix<float,3,3,RowMajor>
struct Vec3f { float v[3]; };

static inline Mat3f mat3_identity() { Mat3f r = {{1,0,0,0,1,0,0,0,1}}; return r; }

// Eigen compute_inverse<Matrix3f,3>: cofactor expansion along column 0, times 1/det.
static inline Mat3f mat3_inverse_eigen(const Mat3f& a)
{
    const float* m = a.m;
#define M(i,j) m[(i)*3+(j)]
#define COF(i,j) (M(((i)+1)%3,((j)+1)%3) * M(((i)+2)%3,((j)+2)%3) - M(((i)+1)%3,((j)+2)%3) * M(((i)+2)%3,((j)+1)%3))
    float c00 = COF(0,0), c10 = COF(1,0), c20 = COF(2,0);
    float det = (c00 * M(0,0) + c10 * M(1,0)) + c20 * M(2,0);
    float invdet = 1.0f / det;
    Mat3f r;
    r.m[0] = c00 * invdet;      r.m[1] = c10 * invdet;      r.m[2] = c20 * invdet;
    r.m[3] = COF(0,1) * invdet; r.m[4] = COF(1,1) * invdet; r.m[5] = COF(2,1) * invdet;
    r.m[6] = COF(0,2) * invdet; r.m[7] = COF(1,2) * invdet; r.m[8] = COF(2,2) * invdet;
#undef COF
#undef M
    return r;
}

static inline Mat3f mat3_mul(const Mat3f& a, const Mat3f& b)
{
    Mat3f r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            r.m[i*3+j] = (a.m[i*3+0]*b.m[0*3+j] + a.m[i*3+1]*b.m[1*3+j]) + a.m[i*3+2]*b.m[2*3+j];
    return r;
}

static inline Vec3f mat3_mulv(const Mat3f& a, const Vec3f& x)
{
    Vec3f r;
    for (int i = 0; i < 3; ++i)
        r.v[i] = (a.m[i*3+0]*x.v[0] + a.m[i*3+1]*x.v[1]) + a.m[i*3+2]*x.v[2];
    return r;
}

static inline Mat3f mat3_transpose(const Mat3f& a)
{
    Mat3f r;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i*3+j] = a.m[j*3+i];
    return r;
}

// x = A^-1 b for symmetric 6x6 A (row-major), Eigen::LDLT style: Bunch-Kaufman-free
// diagonal pivoting (largest |diagonal| first), unit-lower L, solve P^T L^-T D^-1 L^-1 P b.
static inline void ldlt6_solve(const double* A, const double* b, double* x)
{
    const int n = 6;
    double a[36];
    int perm[6];
    std::memcpy(a, A, sizeof(a));
    for (int i = 0; i < n; ++i) perm[i] = i;
    for (int k = 0; k < n; ++k) {
        int piv = k; double best = std::fabs(a[k*n+k]);
        for (int i = k + 1; i < n; ++i) { double v = std::fabs(a[i*n+i]); if (v > best) { best = v; piv = i; } }
        if (piv != k) {   // symmetric row/column swap
            for (int j = 0; j < n; ++j) { double t = a[k*n+j]; a[k*n+j] = a[piv*n+j]; a[piv*n+j] = t; }
            for (int i = 0; i < n; ++i) { double t = a[i*n+k]; a[i*n+k] = a[i*n+piv]; a[i*n+piv] = t; }
            int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
        }
        double d = a[k*n+k];
        if (d == 0.0) continue;
        for (int i = k + 1; i < n; ++i) a[i*n+k] /= d;            // L column
        for (int i = k + 1; i < n; ++i)
            for (int j = k + 1; j <= i; ++j) {
                a[i*n+j] -= a[i*n+k] * d * a[j*n+k];
                a[j*n+i] = a[i*n+j];
            }
    }
    double y[6];
    for (int i = 0; i < n; ++i) y[i] = b[perm[i]];
    for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j) y[i] -= a[i*n+j] * y[j];      // L^-1
    for (int i = 0; i < n; ++i) { double d = a[i*n+i]; y[i] = (std::fabs(d) > DBL_MIN) ? y[i] / d : 0.0; }
    for (int i = n - 1; i >= 0; --i) for (int j = i + 1; j < n; ++j) y[i] -= a[j*n+i] * y[j]; // L^-T
    for (int i = 0; i < n; ++i) x[perm[i]] = y[i];
}

// cv::Rodrigues(rvec(3x1,64F) -> R(3x3,64F)), OpenCV 2.4.9 cvRodrigues2 vector branch.
static inline void rodrigues_cv(const double* r, double* R)
{
    double rx = r[0], ry = r[1], rz = r[2];
    double theta = std::sqrt(rx*rx + ry*ry + rz*rz);
    if (theta < DBL_EPSILON) {
        for (int k = 0; k < 9; ++k) R[k] = (k % 4 == 0) ? 1.0 : 0.0;
        return;
    }
    const double I[9] = {1,0,0,0,1,0,0,0,1};
    double c = std::cos(theta), s = std::sin(theta), c1 = 1.0 - c;
    double itheta = theta ? 1.0 / theta : 0.0;
    rx *= itheta; ry *= itheta; rz *= itheta;
    double rrt[9] = {rx*rx, rx*ry, rx*rz, rx*ry, ry*ry, ry*rz, rx*rz, ry*rz, rz*rz};
    double rx_[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
    for (int k = 0; k < 9; ++k) R[k] = c * I[k] + c1 * rrt[k] + s * rx_[k];
}

// OdometryProvider::computeProjectiveMatrix (OdometryProvider.h:54-68): ksi=[t; w] -> 4x4.
static inline void projective_matrix(const double* ksi, double* Rt)
{
    double R[9];
    rodrigues_cv(ksi + 3, R);
    for (int k = 0; k < 16; ++k) Rt[k] = (k % 5 == 0) ? 1.0 : 0.0;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rt[i*4+j] = R[i*3+j];
    Rt[3] = ksi[0]; Rt[7] = ksi[1]; Rt[11] = ksi[2];
}

static inline void mat4d_mul(const double* a, const double* b, double* c)   // c = a*b, no aliasing
{
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += a[i*4+k] * b[k*4+j];
            c[i*4+j] = s;
        }
}

// Pose update of ICPOdometry.cpp:146-178 / RGBDOdometry.cpp:337-369:
// [Rcurr|tcurr] = [Rprev|tprev] * ([rot|trans])^-1 in float (Eigen::Isometry3f semantics).
static inline void compose_prev_with_inverse(const Mat3f& Rprev, const Vec3f& tprev, const double* resultRt,
                                             Mat3f* Rcurr, Vec3f* tcurr)
{
    Mat3f rot; Vec3f tr;
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) rot.m[i*3+j] = (float)resultRt[i*4+j]; tr.v[i] = (float)resultRt[i*4+3]; }
    Mat3f rinv = mat3_transpose(rot);             // Isometry inverse: R^T, -R^T t
    Vec3f tinv = mat3_mulv(rinv, tr);
    tinv.v[0] = -tinv.v[0]; tinv.v[1] = -tinv.v[1]; tinv.v[2] = -tinv.v[2];
    *Rcurr = mat3_mul(Rprev, rinv);
    Vec3f rt = mat3_mulv(Rprev, tinv);
    tcurr->v[0] = rt.v[0] + tprev.v[0]; tcurr->v[1] = rt.v[1] + tprev.v[1]; tcurr->v[2] = rt.v[2] + tprev.v[2];
}

// Inverse of a rigid 4x4 (double). The reference uses cv::Mat::inv(DECOMP_SVD) on resultRt
// (RGBDOdometry.cpp:211); for a rigid transform the pseudo-inverse equals [R^T | -R^T t] to rounding.
static inline void rigid4d_inverse(const double* T, double* Ti)
{
    for (int k = 0; k < 16; ++k) Ti[k] = (k % 5 == 0) ? 1.0 : 0.0;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Ti[i*4+j] = T[j*4+i];
    for (int i = 0; i < 3; ++i) Ti[i*4+3] = -(Ti[i*4+0]*T[3] + Ti[i*4+1]*T[7] + Ti[i*4+2]*T[11]);
}

} // namespace kto
