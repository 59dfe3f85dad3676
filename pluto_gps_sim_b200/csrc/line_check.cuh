// line_check.cuh — "one straight line per tile" model of the two NCOs and its exact safety check.
//
// The reference advances both NCOs with one binary64 rounding per sample
// (plutogpssim.c:2709 code, 2741 carrier).  What the sample loop CONSUMES, however, is only
//     carrier table index  floor(carr_phase * 512)      plutogpssim.c:2697
//     chip index           (int) code_phase             plutogpssim.c:2737
//     code-period wraps    code_phase >= 1023           plutogpssim.c:2710-2733 (NAV bit counters)
// i.e. floors of the phases.  Every rounding moves a phase by at most half an ulp of the
// binade it lands in, so after n steps from an exactly known state x0 the true phase lies within
// n * (half ulp of the top binade) of the straight line x0 + n*d evaluated in exact integer
// arithmetic.  Scaled to 64-bit fixed point:
//     carrier  F = phase * 2^64 (mod 2^64: the wrap to [0,1) is free), index = F >> 55,
//              per-step error <= 2^-53 cycles = 2^11 units (sums in [1,2) round to 2^-52)
//     code     G = (chips + 1023 * wraps) * 2^47 (never wrapped inside a tile), index = G >> 47,
//              per-step error <= 2^-44 chips = 2^3 units (results < 1024)
// The floors of line and truth agree for a sample unless the line passes within that error bound
// of an index boundary (a multiple of 2^55 resp. 2^47).  Whether ANY of N consecutive samples of
// a line comes that close is a question about min/max of (a + n*d) mod 2^B over n < N, which a
// Euclid-like descent answers in O(log N) steps (minmod below).  k_line_anchor runs that check
// for every (epoch, slot, NCO) per chunk of 32 tiles, then per tile for flagged chunks; the
// (tile, slot) pairs it cannot clear -- about 2^-22 of them -- are walked with the literal
// recurrence by k_line_patch and any differing samples are patched after the main kernel.
// So k_synth_line never needs a segment list: exactness rests on the bound + the check.
#pragma once
#include <stdint.h>

#include "nco_scan.cuh"

namespace gpsiq {

constexpr int LN_FBITS = 55;                 // carrier: index = F >> 55 (9 index bits)
constexpr int LN_GBITS = 47;                 // code:    index = G >> 47 (17 index bits: chip + table placement)
constexpr int64_t LN_EPS_F = 2048 + 2;       // per-step bound, carrier (rounding 2^11 + truncated step 1 + slack)
constexpr int64_t LN_EPS_G = 16 + 2;         // per-step bound, code (rounding 2^3, doubled, + truncated step + slack)

// deviation bound after n steps from an anchor that is itself floor()ed (< 1 unit)
GPSIQ_HD int64_t ln_eps(int is_carrier, int64_t n) { return n * (is_carrier ? LN_EPS_F : LN_EPS_G) + 2; }

// ---- binary64 -> fixed point ----------------------------------------------------
// floor(|x| * 2^scale) for finite x (0 for denormals); saturates at 2^64-1.
GPSIQ_HD uint64_t ln_fixed_abs(double x, int scale) {
    const int64_t b = f64_bits(x) & 0x7fffffffffffffffLL;
    const int e = (int) (b >> 52);
    if (e <= 0) return 0;
    const uint64_t m = ((uint64_t) b & 0xfffffffffffffULL) | (1ULL << 52);
    const int sh = e - 1075 + scale;
    if (sh >= 12) return ~0ULL;
    if (sh >= 0) return m << sh;
    return (sh > -64) ? (m >> (-sh)) : 0;
}
// carrier phase in [0,1] -> F (phase == 1.0, which plutogpssim.c:2745-2746 can round to, saturates: index 511)
GPSIQ_HD uint64_t ln_carr_fixed(double ph) { return ln_fixed_abs(ph, 64); }
GPSIQ_HD uint64_t ln_code_fixed(double chips) { return ln_fixed_abs(chips, LN_GBITS); }
// per-sample line slopes (truncated toward zero: error < 1 unit per step, inside LN_EPS_*)
GPSIQ_HD uint64_t ln_carr_slope(double d) {
    const uint64_t a = ln_fixed_abs(d, 64);
    return (d < 0.0) ? (uint64_t) 0 - a : a;
}
GPSIQ_HD uint64_t ln_code_slope(double d) { return ln_fixed_abs(d, LN_GBITS); }

GPSIQ_HD int ln_ctz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long) v) - 1;
#else
    return __builtin_ctzll(v);
#endif
}

// ---- min over x in [0, n) of (b + a*x) mod m --------------------------------------
// 0 <= b < m, 0 <= a < m, n >= 1.  Returns the exact minimum, except that it may return early
// with ANY attained value < stop (callers only ask "is the minimum below stop?").
// Descent: the minimum of an ascending ramp sequence is its start or the first value after one
// of its K wraps; those K values are themselves an arithmetic sequence modulo the step.  A step
// above m/2 is handled as a descending sequence of step m - a, whose minimum is the last value or
// one of the K troughs before a wrap, again an arithmetic sequence modulo the step.  The modulus
// at least halves every two rounds, and n shrinks at least as fast.
GPSIQ_HD uint64_t minmod(uint64_t b, uint64_t a, uint64_t m, uint64_t n, uint64_t stop) {
    typedef unsigned __int128 u128;
    uint64_t best = b;
    for (int guard = 0; guard < 200; guard++) {
        if (b < best) best = b;
        if (best < stop || a == 0 || n <= 1) return best;
        if (a <= m - a) {
            // ascending: values b + a*x - k*m; K = floor((b + (n-1)*a) / m) wraps
            uint64_t K;
            const u128 tot = (u128) (n - 1) * a + b;
            if ((uint64_t) (tot >> 64) == 0) K = (uint64_t) tot / m;   // the usual case after the first round
            else if ((m & (m - 1)) == 0) K = (uint64_t) (tot >> ln_ctz64(m));  // first round: m = 2^B
            else K = (uint64_t) (tot / m);
            if (K == 0) return best;
            const uint64_t r = m % a;
            const uint64_t nb = (a - ((m - b) % a)) % a;  // (b - m) mod a
            const uint64_t na = (a - r) % a;              // (-m) mod a
            b = nb; m = a; a = na; n = K;
        } else {
            // descending by c = m - a: values b - c*x + k*m
            const uint64_t c = m - a;
            const u128 tot = (u128) (n - 1) * c;
            if (tot <= b) {  // never wraps: the last value is the minimum
                const uint64_t v = b - (uint64_t) tot;
                return v < best ? v : best;
            }
            const u128 need = tot - b;                           // > 0
            uint64_t K;                                          // wraps = ceil(need / m)
            const u128 num = need + (m - 1);
            if ((uint64_t) (num >> 64) == 0) K = (uint64_t) num / m;
            else if ((m & (m - 1)) == 0) K = (uint64_t) (num >> ln_ctz64(m));
            else K = (uint64_t) (num / m);
            const uint64_t fin = (uint64_t) ((u128) K * m - need);  // last value, in [0, m)
            if (fin < best) best = fin;
            b = b % c; a = m % c; m = c; n = K;                  // troughs (b + k*m) mod c, k = 0..K-1
        }
    }
    return 0;  // not reachable (the modulus halves every two rounds); "hazard" is the safe answer
}

// Does any n in [0, N) put a multiple of 2^B inside (A + n*d + lo, A + n*d + hi]  (lo <= hi)?
// A, d are taken modulo 2^B.  This is the hazard test: line values A + n*d, truth and the values
// the main kernel uses all within [lo, hi] of the line => their indices (>> B) agree unless this
// returns true.
GPSIQ_HD bool line_hazard(uint64_t A, uint64_t d, int B, uint64_t N, int64_t lo, int64_t hi) {
    const uint64_t m = 1ULL << B, mask = m - 1;
    if (N == 0) return false;
    const uint64_t w = (uint64_t) (hi - lo);
    if (w >= m) return true;
    if (w == 0) return false;
    const uint64_t A2 = (A + (uint64_t) lo) & mask;
    // max_n (A2 + n d) mod m + w >= m   <=>   min_n ((m-1-A2) + n (m-d)) mod m < w
    return minmod(m - 1 - A2, (m - (d & mask)) & mask, m, N, w) < w;
}

}  // namespace gpsiq
