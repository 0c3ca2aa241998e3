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

} // namespace kt
