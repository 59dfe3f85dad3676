// nco_scan.cuh — exact fast-forward of the reference's binary64 NCO recurrences.
//
// The reference advances both NCOs once per sample with one IEEE-754 rounding
// per step and no closed form:
//     code    : code_phase += f_code*delt; if (>= 1023) -= 1023      plutogpssim.c:2709-2713
//     carrier : carr_phase += f_carr*delt; if (>= 1) -= 1; else if (< 0) += 1
//                                                                    plutogpssim.c:2741-2746
// Bit-exact output needs the exact state at every sample, but a GPU wants to
// start thousands of sample tiles at once.  The bridge is this scan: it
// reproduces n sequential steps in O(#binades crossed) work.
//
// Lemma.  Let x be a binary64 in the binade [2^b, 2^(b+1)),
// ulp g = 2^(b-52).  If RN(x+d) stays in the same binade, then
// RN(x+d) = x + D*g with D = RN(d/g) an integer that does not depend on x
// (x is a multiple of g); on an exact tie RN picks the even mantissa, after
// which the increment is constant as well.  Within a binade the raw IEEE bit
// pattern of a positive double is an affine function of its value, so k such
// steps are ONE integer multiply-add on the bit pattern:
//     bits(x_k) = bits(x_0) + k*D.
// In other words: inside a binade the reference's floating-point NCO *is* a
// fixed-point NCO.  Binade crossings and wraps are done with true additions.
// The subtractions x-1023.0 (x in [1023,1024.4)) and x-1.0 (x in [1,2)) are
// exact; x+1.0 for a small negative x rounds to the 2^-53 grid.
//
// This header is shared by the CUDA kernels (device) and by the host-side unit
// test build (tests/ compile it with g++ and compare against the literal
// per-sample recurrence).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GPSIQ_HD __host__ __device__ __forceinline__
#else
#define GPSIQ_HD inline
#endif

namespace gpsiq {

GPSIQ_HD int64_t f64_bits(double x) {
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(x);
#else
    union { double d; int64_t i; } u; u.d = x; return u.i;
#endif
}
GPSIQ_HD double bits_f64(int64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(b);
#else
    union { double d; int64_t i; } u; u.i = b; return u.d;
#endif
}
// One rounding, never contracted into an FMA.
GPSIQ_HD double add_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b; return r;
#endif
}

constexpr int NCO_CODE = 0;     // wrap at 1023, step > 0, counts wraps
constexpr int NCO_CARRIER = 1;  // wrap at 1 / 0, step of either sign

constexpr int64_t BITS_1023 = 0x408FF80000000000LL;  // bits(1023.0)
constexpr int64_t BITS_1 = 0x3FF0000000000000LL;     // bits(1.0)

// One literal step of the reference recurrence. Returns true if it wrapped.
template <int MODE>
GPSIQ_HD bool nco_step(double& x, double d, int& wraps) {
    double y = add_rn(x, d);
    bool w = false;
    if (MODE == NCO_CODE) {
        if (y >= 1023.0) { y = add_rn(y, -1023.0); wraps++; w = true; }
    } else {
        if (y >= 1.0) { y = add_rn(y, -1.0); w = true; }
        else if (y < 0.0) { y = add_rn(y, 1.0); w = true; }
    }
    x = y;
    return w;
}

// Floor of a/b for 0 <= a < 2^53, 0 < b < 2^53 without a 64-bit integer
// division on the device: both are exact doubles, the round-toward-zero
// quotient is the largest double <= a/b, and every integer below 2^53 is
// representable, so its floor is the exact integer quotient.
GPSIQ_HD int64_t floor_div_pos(int64_t a, int64_t b) {
#if defined(__CUDA_ARCH__)
    return __double2ll_rz(__ddiv_rz(__ll2double_rn(a), __ll2double_rn(b)));
#else
    return a / b;
#endif
}

// Advance x by `count` steps of the MODE recurrence with step d; `wraps`
// accumulates code-period wraps (NCO_CODE).  Exactly equivalent to calling
// nco_step `count` times.
template <int MODE>
GPSIQ_HD void nco_advance(double& x, double d, int count, int& wraps) {
    while (count > 0) {
        double x0 = x;
        bool w = nco_step<MODE>(x, d, wraps);
        count--;
        if (w || count < 3) continue;
        int64_t b0 = f64_bits(x0), b1 = f64_bits(x);
        if ((b0 >> 52) != (b1 >> 52)) continue;  // left the binade
        // second literal step: after it the increment is tie-free and constant
        w = nco_step<MODE>(x, d, wraps);
        count--;
        if (w) continue;
        int64_t b2 = f64_bits(x);
        if ((b2 >> 52) != (b1 >> 52)) continue;
        int64_t delta = b2 - b1;
        if (delta == 0) return;  // step below half an ulp: the phase no longer moves
        int64_t k;
        if (delta > 0) {
            int64_t top = ((b2 >> 52) + 1) << 52;  // first pattern of the next binade
            const int64_t lim = (MODE == NCO_CODE) ? BITS_1023 : BITS_1;
            if (lim < top) top = lim;
            k = floor_div_pos(top - 1 - b2, delta);  // results stay <= top-1
        } else {
            int64_t bot = (b2 >> 52) << 52;  // first pattern of this binade
            k = floor_div_pos(b2 - bot - 1, -delta);  // results stay >= bot+1
        }
        if (k > count) k = count;
        x = bits_f64(b2 + k * delta);
        count -= (int) k;
    }
}

}  // namespace gpsiq
