// kintinuous_b200 -- exact fast-forward of a float running sum.
//
// The reference's tsdf23 advances v_x, v_y along z by REPEATED float additions of a per-launch constant (tsdf_volume.cu:565-574); the rounding
// of those running sums decides which depth pixel a voxel reads, so a kernel that starts a column at z = zlo must hold exactly the value
// the reference reaches after zlo additions.  Replaying them one by one costs O(zlo) dependent FADDs per thread and z-chunk (about a
// quarter of integrate_kernel's issued instructions).  replay_add(x, a, k) returns the SAME bits in O(number of binades crossed):
//   while the sum stays inside one binade (ulp u = 2^(e-23)) every addition moves the mantissa by the same integer A = round(|a| / u)
//   -- x + a = (X +- A) u -+ r with |r| < u/2 rounds to (X +- A) u -- so n steps are X +- n A in integer arithmetic; a step that
//   leaves the binade, a tie (|r| = u/2 exactly: round-half-even alternates) or an x smaller than a is executed as a real addition.
// Round-to-nearest-even; denormal inputs are taken as zero like the kernels' --ftz=true arithmetic (results never become denormal
// on the fast path).  Compiles for the host too: tests/test_replay_add_host.py checks it against the plain loop.
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#ifndef KT_HD
#define KT_HD __host__ __device__
#endif

namespace kt {

KT_HD __forceinline__ uint32_t rp_bits(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
KT_HD __forceinline__ float rp_float(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
KT_HD __forceinline__ float rp_add(float x, float a) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(x, a);
#else
    volatile float r = x + a; return r;
#endif
}

KT_HD inline float replay_add(float x, float a, int k)
{
    const uint32_t ab = rp_bits(a);
    const int ea = (int)((ab >> 23) & 0xffu);
    if (ea == 0) return k > 0 ? rp_add(x, rp_float(ab & 0x80000000u)) : x;   // a == +-0 (or denormal, flushed to it): one addition settles it (-0 + +0 = +0)
    const uint32_t Ma = (ab & 0x7fffffu) | 0x800000u;          // 24-bit significand of |a|
    while (k > 0) {
        const uint32_t xb = rp_bits(x);
        const int ex = (int)((xb >> 23) & 0xffu);
        const int sh = ex - ea;                                // u(x) = 2^sh * u(a)
        const uint32_t X = (xb & 0x7fffffu) | 0x800000u;
        const bool same = ((xb ^ ab) >> 31) == 0u;
        // the closed form needs: x normal, |a| not above x's binade, and -- when the magnitude shrinks -- x not sitting on the binade's lower
        // edge (from there the very first step already lands in the binade below, where the ulp is half as large)
        if (ex != 0 && ex != 255 && ea != 255 && sh >= 0 && (same || X != 0x800000u)) {
            if (sh > 24) return x;                             // |a| < u/2: x + a == x, now and for every further step
            uint32_t A, rem, half;
            if (sh == 0) { A = Ma; rem = 0; half = 1; }
            else { A = Ma >> sh; rem = Ma & ((1u << sh) - 1u); half = 1u << (sh - 1); if (rem > half) ++A; }
            if (rem != half || sh == 0) {                      // not a tie
                if (A == 0) return x;                          // |a| < u/2 (sh == 24, rem < half)
                // stay strictly inside the binade: (X - nA) u - r must not drop below 2^23 u, where the ulp halves
                const uint32_t room = same ? (0xffffffu - X) : (X - 0x800001u);
                uint32_t n = room / A;
                if (n > (uint32_t)k) n = (uint32_t)k;
                if (n > 0) {
                    const uint32_t Xn = same ? X + n * A : X - n * A;
                    x = rp_float((xb & 0xff800000u) | (Xn & 0x7fffffu));
                    k -= (int)n;
                    if (k == 0) break;
                }
            }
        }
        {   // the step that leaves the binade / a tie / |x| < |a| / zero / inf / nan
            const float xn = rp_add(x, a);
            if (rp_bits(xn) == rp_bits(x)) return x;           // a fixed point stays one
            x = xn;
            --k;
        }
    }
    return x;
}


// The reference's tsdf23 as COMPILED does not add a float increment: nvcc contracts  v_x += Rcurr_inv.z * cell_size.z * intr.fx  into
//   v_x = fma(R_z * cell, f, v_x)          (SASS of the reference build: FMUL R4 = cell * R_z once, FFMA R9 = R4 * fx + R9 per z step)
// i.e. the addend is the EXACT 48-bit product P = m * f, not its float rounding.  The two differ only when a running sum sits within a
// fraction of an ulp of a rounding boundary -- a handful of voxels per million pick the neighbouring depth pixel -- which is exactly what a
// bit-exact replay against the reference at 512^3 exposed.  replay_fma(x, m, f, k) returns the bits of k steps x <- fma(m, f, x): same closed
// form as replay_add with P kept exact in double (24 x 24 bits fit 53), A = round(P / u) and the remainder compared with u/2 exactly.
KT_HD __forceinline__ float rp_fma(float m, float f, float x) {
#ifdef __CUDA_ARCH__
    return __fmaf_rn(m, f, x);
#else
    return fmaf(m, f, x);
#endif
}

// 2^e as a double, e in [-1022, 1023], from its exponent bits
KT_HD __forceinline__ double rp_pow2(int e) {
    const uint64_t b = (uint64_t)(1023 + e) << 52;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)b);
#else
    double d; memcpy(&d, &b, 8); return d;
#endif
}

KT_HD inline float replay_fma(float x, float m, float f, int k)
{
    if (k <= 0) return x;
    const double P = (double)m * (double)f;                     // exact
    if (P == 0.0 || !(P == P) || P - P != 0.0) {                // zero / nan / inf addend: a real step decides, and it is a fixed point afterwards
        const float x1 = rp_fma(m, f, x);
        return (k == 1 || rp_bits(rp_fma(m, f, x1)) == rp_bits(x1)) ? x1 : rp_fma(m, f, x1);
    }
    while (k > 0) {
        const uint32_t xb = rp_bits(x);
        const int ex = (int)((xb >> 23) & 0xffu);
        if (ex != 0 && ex != 255) {
            // q = P / u scaled into the magnitude direction of x: positive q_m grows |x|.  u = 2^(ex - 150).
            if (ex > 150 - 62 && ex < 150 + 62) {
                const double scale = rp_pow2(150 - ex);          // 1 / u as an exact power of two, built from its exponent bits (no division)
                double q = P * scale;                           // exact (power-of-two scaling, no overflow in these ranges)
                if (xb >> 31) q = -q;
                if (q < 16777216.0 && q > -16777216.0) {
                    const double A = rint(q);                   // nearest integer, ties to even (irrelevant: ties are excluded below)
                    const double r = q - A;                     // exact
                    if (r != 0.5 && r != -0.5) {
                        const uint32_t X = (xb & 0x7fffffu) | 0x800000u;
                        if (A == 0.0) {
                            // |P| < u/2: x is a fixed point -- unless x sits on its binade's lower edge and shrinks (finer ulp below)
                            if (q >= 0.0 || X != 0x800000u) return x;
                        } else {
                            // |A| < 2^24 and the room inside the binade < 2^24: 32-bit unsigned division (the 64-bit one is a long emulated
                            // sequence on the GPU, and this chain is the set-up latency of every integrate_kernel thread)
                            const int32_t Ai = (int32_t)A;
                            uint32_t n;
                            if (Ai > 0) n = (0xffffffu - X) / (uint32_t)Ai;
                            else n = X > 0x800000u ? (X - 0x800001u) / (uint32_t)(-Ai) : 0u;
                            if (n > (uint32_t)k) n = (uint32_t)k;
                            if (n > 0) {
                                const uint32_t Xn = (uint32_t)((int32_t)X + (int32_t)n * Ai);
                                x = rp_float((xb & 0xff800000u) | (Xn & 0x7fffffu));
                                k -= (int)n;
                            }
                        }
                    }
                }
            }
        }
        if (k == 0) break;
        {   // the step that leaves the binade / a tie / |x| below the addend / zero / inf / nan
            const float xn = rp_fma(m, f, x);
            if (rp_bits(xn) == rp_bits(x)) return x;
            x = xn;
            --k;
        }
    }
    return x;
}

} // namespace kt
