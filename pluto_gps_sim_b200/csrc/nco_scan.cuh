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
// (x is a multiple of g); on an exact tie RN picks the even mantissa, so from
// an even mantissa the increment is constant as well.  Within a binade the raw
// IEEE bit pattern of a positive double is an affine function of its value, so
// k such steps are ONE integer multiply-add on the bit pattern:
//     bits(x_k) = bits(x_0) + k*D.
// In other words: inside a binade the reference's floating-point NCO *is* a
// fixed-point NCO.  Binade crossings and wraps are done with true additions.
// The subtractions x-1023.0 (x in [1023,1024.4)) and x-1.0 (x in [1,2)) are
// exact; x+1.0 for a small negative x rounds to the 2^-53 grid.
//
// The per-binade increment D_b is the step's own significand shifted to the
// binade's ulp and rounded (binade_delta): two registers (StepInfo), no table,
// no memory traffic.  Every scan steps segment by segment:
// [k in-binade steps by one multiply-add] [one true step].
//
// This header is shared by the CUDA kernels (device) and by the host-side
// entry points gpsiq_nco_advance / gpsiq_carrier_chain_host, which the tests
// compare against the literal per-sample recurrence.
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
GPSIQ_HD double pow2_of(int64_t biased_exp) { return bits_f64(biased_exp << 52); }

constexpr int NCO_CODE = 0;     // wrap at 1023, step > 0, counts wraps
constexpr int NCO_CARRIER = 1;  // wrap at 1 / 0, step of either sign

constexpr int64_t BITS_1023 = 0x408FF80000000000LL;  // bits(1023.0)
constexpr int64_t BITS_1 = 0x3FF0000000000000LL;     // bits(1.0)
constexpr int NBINADE = 16;

template <int MODE> struct NcoTraits;
template <> struct NcoTraits<NCO_CODE> { static constexpr int64_t ETOP = 1032; static constexpr int64_t LIM = BITS_1023; };
template <> struct NcoTraits<NCO_CARRIER> { static constexpr int64_t ETOP = 1022; static constexpr int64_t LIM = BITS_1; };

// What a scan keeps per chain about its step: the step itself and ~1/|d| (single precision: only ever used for a
// first guess that exact integer arithmetic then corrects).
struct StepInfo {
    double d;
    float rd;
};

GPSIQ_HD float fast_rcp(float v) {
#if defined(__CUDA_ARCH__)
    return __frcp_rn(v);
#else
    return 1.0f / v;
#endif
}
GPSIQ_HD int64_t round_even_ll(double t) {  // |t| < 2^62; nearest integer, ties to even
#if defined(__CUDA_ARCH__)
    return __double2ll_rn(t);
#else
    return (int64_t) __builtin_rint(t);  // default rounding mode: to nearest even
#endif
}
GPSIQ_HD float pow2f_of(int biased_exp) {  // 2^(biased_exp - 127), 1 <= biased_exp <= 254
#if defined(__CUDA_ARCH__)
    return __int_as_float(biased_exp << 23);
#else
    union { float f; int32_t i; } u; u.i = biased_exp << 23; return u.f;
#endif
}
GPSIQ_HD double abs_f64(double t) { return bits_f64(f64_bits(t) & 0x7fffffffffffffffLL); }

GPSIQ_HD StepInfo step_info(double d) {
    StepInfo s;
    s.d = d;
    const float a = (float) abs_f64(d);
    s.rd = (a > 0.f) ? fast_rcp(a) : 0.f;
    return s;
}

// In-binade increment of the bit pattern for a state with biased exponent ex (64 <= ex), from an EVEN mantissa (any
// mantissa if the binade has no tie): D = RN(d / ulp_x), ulp_x = 2^(ex - 1075).  The division by the ulp is a
// power-of-two scaling, exact in binary64, and round-to-nearest-even of the quotient is exactly what the reference's
// addition does to the sum's mantissa (x even: ties go to the even D).  tie: d / ulp_x is an exact half-integer --
// an odd mantissa then steps differently once.  false: no usable increment (|d| >= the binade, or not finite).
GPSIQ_HD bool binade_delta(const StepInfo& si, int64_t ex, int64_t& delta, bool& tie) {
    tie = false;
    if (ex < 64 || ex > 2046) return false;
    const double t = si.d * pow2_of(2098 - ex);  // d / ulp_x
    if (!(abs_f64(t) < 0x1p52)) return false;
    delta = round_even_ll(t);
    tie = abs_f64(t - (double) delta) == 0.5;
    return true;
}

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

// Take up to maxk steps that provably stay inside x's binade (and below the
// wrap limit) with one multiply-add on the bit pattern.  Returns the number of
// steps taken (possibly 0); none of them wraps.  edge_gap (if not null)
// receives the distance, in ulps of the binade, between the last iterate and
// the binade edge / wrap limit it is moving toward (>= 1).
template <int MODE>
GPSIQ_HD int run_in_binade(double& x, const StepInfo& si, int maxk, int64_t* edge_gap = nullptr, int64_t* delta_out = nullptr) {
    const int64_t b = f64_bits(x);
    const int64_t e = b >> 52;
    if (edge_gap) *edge_gap = 0;
    if (e < 960 || e > NcoTraits<MODE>::ETOP || maxk <= 0) return 0;  // zero / tiny (< 2^-63: true steps) / negative / above the wrap limit
    int64_t delta;
    bool tie;
    if (!binade_delta(si, e, delta, tie)) return 0;
    if (tie && (b & 1)) return 0;  // odd mantissa in a tie binade: take a true step first
    if (delta_out) *delta_out = delta;
    if (delta == 0) return maxk;  // |d| < ulp/2: the state no longer moves
    int64_t num, ad;
    if (delta > 0) {
        int64_t top = (e + 1) << 52;
        if (NcoTraits<MODE>::LIM < top) top = NcoTraits<MODE>::LIM;
        num = top - 1 - b;  // iterates must stay <= top-1
        ad = delta;
    } else {
        num = b - (e << 52) - 1;  // iterates must stay >= first pattern of the binade + 1
        ad = -delta;
    }
    if (num < ad) return 0;
    // k = min(floor(num/ad), maxk) without an integer or FP64 division
    int64_t k;
    // first guess of num / ad: 1/ad = (1/|d|) * ulp_x up to rounding -- a power-of-two scaling of the chain's 1/|d|
    const float qf = (float) num * (si.rd * pow2f_of((int) e - 1075 + 127));
    if (qf >= (float) maxk + 4.0f) {
        k = maxk;
    } else {
        k = (int64_t) qf;
        int64_t r = num - k * ad;
        while (r < 0) { k--; r += ad; }
        while (r >= ad) { k++; r -= ad; }
        if (k > maxk) k = maxk;
    }
    x = bits_f64(b + k * delta);
    if (edge_gap) *edge_gap = num - k * ad + 1;
    return (int) k;
}

// Advance x by `count` steps of the MODE recurrence with step d; `wraps`
// accumulates code-period wraps (NCO_CODE).  Exactly equivalent to calling
// nco_step `count` times.
template <int MODE>
GPSIQ_HD void nco_advance(double& x, double d, const StepInfo& tab, int count, int& wraps) {
    while (count > 0) {
        count -= run_in_binade<MODE>(x, tab, count);
        if (count == 0) break;
        nco_step<MODE>(x, d, wraps);
        count--;
    }
}

// ===========================================================================
// Parallel exact carrier scan: speculate -> translate -> verify
// ===========================================================================
// The carrier phase never restarts (plutogpssim.c:2741-2746; only allocation
// re-seeds it, plutogpssim.c:1964), so its exact value at epoch e depends on
// every sample before it: one serial chain per channel across the whole
// stream.  Scanning it segment by segment is latency bound on a GPU.  This
// scheme makes all but O(1 carrier cycle) of each epoch parallel:
//
//  (1) SPECULATE (spec_scan_epoch, one chain per (epoch, channel), all in
//      parallel): scan the epoch from an ESTIMATED start phase.  After the
//      first wrap of the run the state sits on a coarse grid: a wrap leaves a
//      multiple of 2^-52 (step > 0: the sum was rounded in [1,2), and x-1 is
//      exact) or of 2^-53 (step < 0: RN(y+1.0) lands in [0.5,1)).
//  (2) TRANSLATE.  Let x'_n be the speculative states from that first wrap on
//      and x_n the true ones, x_n1 = x'_n1 + D with D a multiple of 2^-52.
//      Every rounding below 1.0 is to a grid of 2^-53 or finer, so D is an EVEN
//      multiple of it and RN(x'+D+d) = RN(x'+d)+D, ties included -- provided
//      every decision (binade of each sum, wrap or not) comes out the same for
//      both runs.  Then x_n = x'_n + D for the rest of the epoch, and that sum
//      is exact in binary64 (same binade, D on its grid).
//      step > 0 needs one exclusion: sums in [1,2) round to 2^-52, of which D
//      may be an odd multiple; a tie there needs step == 0 (mod 2^-53), so such
//      epochs (1 in ~2^9) are not speculated.  step < 0 wraps leave multiples of
//      2^-53: the speculation is run for both parities and the one making D a
//      multiple of 2^-52 is used.
//  (3) VERIFY (chain_epoch, serial per channel, O(one carrier cycle) per
//      epoch): from the exact epoch start run the exact scan up to the first
//      wrap; require the same wrap index as the speculation and |D| below the
//      speculation's decision margin (the smallest distance of any speculative
//      sum to a power of two / zero, minus rounding slack).  If so the epoch's
//      end state is xend' + D; otherwise (or if not speculated) the epoch is
//      scanned serially.  Exactness never depends on the estimate's quality,
//      only the speed does.

// distance of |r| to the nearest power of two (the ends of its binade), minus one ulp of slack
GPSIQ_HD double binade_margin(double r) {
    const int64_t b = f64_bits(r) & 0x7fffffffffffffffLL;
    const int64_t e = b >> 52;
    if (e == 0 || e >= 2046) return -1.0;  // zero / denormal / huge: no margin
    const double ar = bits_f64(b);
    const double lo = pow2_of(e), hi = pow2_of(e + 1);
    const double m = (ar - lo < hi - ar) ? ar - lo : hi - ar;  // both differences are exact
    return m - pow2_of(e - 52);
}

// A positive step that is a multiple of 2^-53 ("tie-capable": 1 epoch in ~2^9) can produce exact ties in [1, 2), where
// a translation by an ODD multiple of 2^-52 does not commute with round-to-nearest-even.  The chunk- and epoch-level
// speculations leave such epochs out (carr_step_speculable); the chains of levels 3-5 scan them serially as part of
// their own speculative trajectory and record the FIRST tie-wrap after the trajectory's first wrap: a true run that is
// that trajectory + D follows it exactly up to the event, and from the event on is the trajectory + D + k * 2^-52 if
// D is an odd multiple of 2^-52 (the two runs round the tie in opposite directions; the new shift is an even multiple,
// which commutes with every later rounding, ties included), and + D throughout if D is an even multiple.
struct TieEvent {
    int pos;   // position (in the trajectory's own units) of the first sample whose state follows the tie-wrap; -1: none
    int k;     // +1: the trajectory rounded the tie down (an odd-shifted run rounds up), -1: the other way round
};
// shift (translation) in force after the event, given the shift D before it
GPSIQ_HD double tie_shift(double D, const TieEvent& t) {
    if (t.pos < 0) return D;
    const double q = D * 0x1p52;                                   // exact: D is a small multiple of 2^-52
    const bool odd = ((long long) q) & 1;
    return odd ? D + (double) t.k * 0x1p-52 : D;                   // exact
}

struct CarrSpec {   // result of one speculative epoch scan
    double xw1;     // state right after the first wrap of the run (sample index n1)
    double xend;    // state after the last sample of the epoch
    double margin;  // decision margin from n1 on; <= 0: unusable
    int n1;         // index of the first sample whose state follows a wrap; N if the run never wraps
    int pad;        // variant 0: 0.  Variant 1 of a chunk run: -1 = scanned for real (plane 1 holds its tile starts);
                    // >= 0 = DERIVED from variant 0 (spec_derive_variant1): ((tie_n + 1) << 2) | (tie_up << 1) | (s < 0)
};

// One carrier step with margin tracking (TRACK) -- same arithmetic as nco_step<NCO_CARRIER>.
// tie (if not null): set to 1 / 2 when a NEGATIVE wrap's y + 1.0 was an exact rounding tie on the 2^-53 grid and was
// rounded down / up (to the even neighbour): the one event at which a run shifted by an odd multiple of 2^-53 stops
// being an exact translate of this one (spec_derive_variant1).
// ptie (if not null): set to +1 / -1 when a POSITIVE wrap's sum x + d was an exact rounding tie on the 2^-52 grid of
// [1, 2) and was rounded down / up (to the even neighbour): the one event at which a run shifted by an ODD multiple of
// 2^-52 stops being an exact translate of this one -- it rounds the other way, and from then on the shift is larger /
// smaller by 2^-52 (even, so it commutes with every later rounding).  Needs d == 0 (mod 2^-53): see TieEvent.
template <bool TRACK>
GPSIQ_HD bool carr_step(double& x, double d, double& margin, int* tie = nullptr, int* ptie = nullptr) {
    double y = add_rn(x, d);
    bool w = false;
    if (TRACK) { const double m = binade_margin(y); if (m < margin) margin = m; }
    if (y >= 1.0) {
        if (TRACK && ptie) {
            const double bb = add_rn(y, -x);                                   // TwoSum: y + t == x + d exactly
            const double t = add_rn(add_rn(x, -add_rn(y, -bb)), add_rn(d, -bb));
            if (t == 0x1p-53) *ptie = 1; else if (t == -0x1p-53) *ptie = -1;   // rounded down / up
        }
        y = add_rn(y, -1.0); w = true;
    }
    else if (y < 0.0) {
        const double z = add_rn(y, 1.0);
        if (TRACK && tie) {
            const double err = add_rn(add_rn(z, -1.0), -y);   // both differences are exact: the rounding error of y + 1.0
            if (err == 0x1p-54) *tie = 2; else if (err == -0x1p-54) *tie = 1;
        }
        y = z; w = true;
        if (TRACK) { const double m = ((1.0 - y < y - 0.5) ? 1.0 - y : y - 0.5) - 0x1p-53; if (m < margin) margin = m; }
    }
    x = y;
    return w;
}

// Advance up to `count` carrier steps; with stop_at_wrap, return right after
// the first step that wrapped.  Returns the number of steps taken.
// pt (if not null, TRACK only): receives the first positive tie-wrap (TieEvent), its position = pos_base + the number of
// steps taken up to and including the tying step.
template <bool TRACK>
GPSIQ_HD int carr_advance(double& x, double d, const StepInfo& tab, int count, bool stop_at_wrap, bool& wrapped,
                          double& margin, TieEvent* pt = nullptr, int pos_base = 0) {
    const int count0 = count;
    wrapped = false;
    while (count > 0) {
        int64_t gap;
        const int k = run_in_binade<NCO_CARRIER>(x, tab, count, TRACK ? &gap : nullptr);
        count -= k;
        if (TRACK && k > 0) {
            // the last iterate of the run is the one closest to the edge it approaches;
            // one ulp of slack for the rounding of its sum
            const double m = (double) (gap - 1) * pow2_of((f64_bits(x) >> 52) - 52);
            if (m < margin) margin = m;
        }
        if (count == 0) break;
        int ptie = 0;
        const bool w = carr_step<TRACK>(x, d, margin, nullptr, (TRACK && pt) ? &ptie : nullptr);
        count--;
        if (TRACK && pt && ptie && pt->pos < 0) { pt->pos = pos_base + (count0 - count); pt->k = ptie; }
        if (w) { wrapped = true; if (stop_at_wrap) break; }
    }
    return count0 - count;
}

// Flat segment walk: one loop iteration = [one in-binade run][one true step], whatever the tiling -- the lanes of a
// warp (neighbouring epochs of one satellite) then execute nearly the same number of iterations.  The state at every
// tile start passed (the state BEFORE that sample) goes to ck[tile * stride]; starts inside a run come from the run's
// closed form on the bit pattern.  n, next: sample indices relative to the scanned range (tile starts = multiples of T
// from its first sample); next = the next tile start not yet emitted (>= n), t = its tile number.
struct CarrWalk { double x; int n; int next; int t; int tie_n; int tie_up; };  // tie_n: sample index right after the
                                                                               // first tie-wrap of the tracked walk (-1: none)

// warp_any: on the device, with the ballot mask of the lanes that walk together, the loop runs in lockstep until the
// LAST lane is done (a lane that is done idles): without it lanes that leave a loop at different iterations never
// meet again (no reconvergence point inside a loop) and the warp degenerates into a dozen groups.  mask 0 / host: plain.
GPSIQ_HD bool warp_any(unsigned mask, bool pred) {
#if defined(__CUDA_ARCH__)
    return mask ? (__any_sync(mask, pred) != 0) : pred;
#else
    (void) mask;
    return pred;
#endif
}

// Walks from w.n to n_end.  Returns true if it stopped right after a wrapping step (stop_at_wrap), false at n_end.
template <bool TRACK>
GPSIQ_HD bool carr_walk(CarrWalk& w, const StepInfo& si, int n_end, int T, bool stop_at_wrap, double* ck, size_t stride,
                        double& margin, unsigned mask = 0) {
    bool stopped = false;
    while (warp_any(mask, w.n < n_end && !stopped)) {
        if (!(w.n < n_end) || stopped) continue;
        if (w.n == w.next) { ck[(size_t) w.t * stride] = w.x; w.t++; w.next += T; }
        int64_t gap, delta = 0;
        const int64_t b0 = f64_bits(w.x);
        const int k = run_in_binade<NCO_CARRIER>(w.x, si, n_end - w.n, TRACK ? &gap : nullptr, &delta);
        if (k > 0) {
            const int n1 = w.n + k;
            while (w.next < n1) {  // tile starts strictly inside the run
                ck[(size_t) w.t * stride] = bits_f64(b0 + (int64_t) (w.next - w.n) * delta);
                w.t++; w.next += T;
            }
            w.n = n1;
            if (TRACK) {
                // the last iterate of the run is the one closest to the edge it approaches;
                // one ulp of slack for the rounding of its sum
                const double m = (double) (gap - 1) * pow2_of((f64_bits(w.x) >> 52) - 52);
                if (m < margin) margin = m;
            }
            if (w.n >= n_end) continue;
            if (w.n == w.next) { ck[(size_t) w.t * stride] = w.x; w.t++; w.next += T; }
        }
        int tie = 0;
        const bool wrapped = carr_step<TRACK>(w.x, si.d, margin, &tie);
        w.n++;
        if (TRACK && tie && w.tie_n < 0) { w.tie_n = w.n; w.tie_up = tie == 2; }
        if (wrapped && stop_at_wrap) stopped = true;
    }
    return stopped;
}

// Predicted systematic rounding drift of the carrier recurrence over n steps of
// step d: inside binade bi every step errs by (D_b*ulp_b - d), and the phase
// spends a fraction 2^-(bi+1) of its steps there.  Only used to improve the
// START-PHASE ESTIMATES of the speculative scans (fewer serial fallbacks);
// never part of an exact result.
GPSIQ_HD double carr_drift_estimate(double d, const StepInfo& tab, int n) {
    double acc = 0.0;
    for (int bi = 0; bi < NBINADE; bi++) {
        int64_t delta;
        bool tie;
        if (!binade_delta(tab, 1022 - bi, delta, tie) || !(abs_f64(d) < pow2_of(1022 - bi - 2))) continue;  // (|d| < a quarter of the binade)
        const double step_b = (double) delta * pow2_of(1022 - bi - 52);  // exact: D_b * ulp_b
        acc += (step_b - d) * pow2_of(1022 - bi);                               // width of [2^-(bi+1), 2^-bi)
    }
    return acc * (double) n;
}

// true if a run with this step can be translated at all (levels 3-5; ties are handled through TieEvent)
GPSIQ_HD bool carr_step_translatable(double d) {
    const int64_t e = (f64_bits(d) & 0x7fffffffffffffffLL) >> 52;
    return e != 0 && e < 1021;  // not zero/denormal, |d| < 0.25
}

// true if an epoch with this step may be speculated (levels 1-2)
GPSIQ_HD bool carr_step_speculable(double d) {
    const int64_t b = f64_bits(d) & 0x7fffffffffffffffLL;
    const int64_t e = b >> 52;
    if (e == 0 || e >= 1021) return false;  // zero/denormal, or |d| >= 0.25
    if (d > 0.0) {
        // exclude d == 0 (mod 2^-53): lowest set bit of |d| at 2^-53 or above
        const int64_t m = (b & 0xfffffffffffffLL) | (1LL << 52);
        const int64_t low = m & -m;  // lowest set bit of the significand
        int tz = 0;
        while ((low >> tz) != 1) tz++;
        if ((int) e - 1075 + tz >= -53) return false;
    }
    return true;
}

// (1) speculative scan of tiles [t0, t1) of an epoch from x, an estimate of the
// phase at sample t0*T.  ck[t*ck_stride] receives the state at the start of tile
// t (absolute tile index).  out.n1 is relative to sample t0*T; it equals the
// range's sample count if the run never wraps.
GPSIQ_HD void spec_scan_range(double x, double d, const StepInfo& tab, int N, int T, int t0, int t1, int variant,
                              double* ck, size_t ck_stride, CarrSpec& out, unsigned mask = 0, int* tie_n = nullptr,
                              int* tie_up = nullptr) {
    const int m_total = ((t1 * T < N) ? t1 * T : N) - t0 * T;
    double margin = 1.0, dummy = 1.0;
    CarrWalk w;
    w.x = x; w.n = 0; w.next = 0; w.t = t0; w.tie_n = -1; w.tie_up = 0;
    out.n1 = m_total;
    out.xw1 = 0.0;
    const bool seen_wrap = carr_walk<false>(w, tab, m_total, T, true, ck, ck_stride, dummy, mask);  // up to the first wrap
    bool bad = false;
    if (seen_wrap) {
        if (variant == 1) w.x = (w.x + 0x1p-53 < 1.0) ? w.x + 0x1p-53 : w.x - 0x1p-53;  // other parity of the 2^-53 grid
        if (!(w.x >= 0.0 && w.x < 1.0)) bad = true;
        out.n1 = w.n;
        out.xw1 = w.x;
    }
    // the rest, margin-tracked (every lane of the mask calls it: a lane that never wrapped is already at the end)
    carr_walk<true>(w, tab, m_total, T, false, ck, ck_stride, margin, mask);
    if (bad) margin = -1.0;
    out.xend = w.x;
    out.margin = (seen_wrap && carr_step_speculable(d)) ? margin : -1.0;
    out.pad = (variant == 1) ? -1 : 0;
    if (tie_n) { *tie_n = w.tie_n; *tie_up = w.tie_up; }
}

// A negative step needs both parities of the 2^-53 grid speculated (the parity of the true post-wrap state is not
// known in advance), but the second run need not be scanned: shifted by s * 2^-53 at the first wrap it stays an exact
// translate of the first one -- every grid below 0.5 is finer (the shift is an even multiple of it), and on the 2^-53
// grid of [0.5, 1) a shift by one grid unit commutes with round-to-nearest except on an exact tie -- until the first
// wrap whose y + 1.0 IS an exact tie: there both runs round to the EVEN neighbour, and from then on they differ by
// 0 or by 2 * 2^-53 (an even shift, which commutes with everything from there on).  Ties of ordinary additions in
// [0.5, 1) need |d| == 2^-54 (mod 2^-53): such steps are scanned for real (carr_step_odd_shift_safe).
GPSIQ_HD bool carr_step_odd_shift_safe(double d) {
    const int64_t b = f64_bits(d) & 0x7fffffffffffffffLL;
    const int64_t e = b >> 52;
    if (e == 0) return false;
    const int64_t m = (b & 0xfffffffffffffLL) | (1LL << 52);
    int tz = 0;
    while (!((m >> tz) & 1)) tz++;
    return (int) e - 1075 + tz != -54;   // the lowest set bit of |d| is not exactly 2^-54
}

// shift of the derived variant-1 run against variant 0, in units of 2^-53, at chunk-relative sample index n
GPSIQ_HD int derived_shift(const CarrSpec& s1, int n) {
    const int tie_n = (s1.pad >> 2) - 1, s = (s1.pad & 1) ? -1 : 1;
    if (n < s1.n1) return 0;                          // before the first wrap the two runs are the same run
    if (tie_n < 0 || n < tie_n) return s;
    const bool up = (s1.pad & 2) != 0;                // the tie was rounded up: 1 + y = (k + 1/2) * 2^-53 with k odd
    return s > 0 ? (up ? 0 : 2) : (up ? -2 : 0);
}

GPSIQ_HD void spec_derive_variant1(const CarrSpec& s0, int m_total, int tie_n, int tie_up, CarrSpec& s1) {
    s1.margin = -1.0; s1.n1 = -1; s1.xw1 = 0.0; s1.xend = 0.0; s1.pad = 0;
    if (!(s0.margin > 0.0) || s0.n1 >= m_total) return;            // never wrapped / unusable: so is its shifted twin
    const int s = (s0.xw1 + 0x1p-53 < 1.0) ? 1 : -1;               // other parity of the 2^-53 grid (as spec_scan_range)
    s1.n1 = s0.n1;
    s1.xw1 = s0.xw1 + (double) s * 0x1p-53;
    s1.pad = ((tie_n + 1) << 2) | (tie_up ? 2 : 0) | (s < 0 ? 1 : 0);
    s1.xend = add_rn(s0.xend, (double) derived_shift(s1, m_total) * 0x1p-53);
    s1.margin = s0.margin - 0x1p-51;                               // the shift itself eats into every decision's margin
    if (!(s1.xw1 >= 0.0 && s1.xw1 < 1.0) || !(s1.xend >= 0.0 && s1.xend < 1.0)) s1.margin = -1.0;
}

// Both parity variants of one chunk: variant 0 scanned, variant 1 derived from it (or scanned, for the rare steps
// whose ordinary additions can tie).  ck0 / ck1: the chunk planes of the two variants.
GPSIQ_HD void spec_scan_chunk(double x, double d, const StepInfo& tab, int N, int T, int t0, int t1, double* ck0, double* ck1,
                              size_t ck_stride, CarrSpec& out0, CarrSpec& out1, unsigned mask = 0) {
    int tie_n = -1, tie_up = 0;
    spec_scan_range(x, d, tab, N, T, t0, t1, 0, ck0, ck_stride, out0, mask, &tie_n, &tie_up);
    out1.margin = -1.0; out1.n1 = -1; out1.xw1 = 0.0; out1.xend = 0.0; out1.pad = 0;
    const bool neg = d < 0.0;
    const bool real = neg && !carr_step_odd_shift_safe(d);
#if defined(__CUDA_ARCH__)
    const unsigned m1 = mask ? __ballot_sync(mask, real) : 0u;
#else
    const unsigned m1 = 0u;
#endif
    if (real) spec_scan_range(x, d, tab, N, T, t0, t1, 1, ck1, ck_stride, out1, m1);
    else if (neg) spec_derive_variant1(out0, ((t1 * T < N) ? t1 * T : N) - t0 * T, tie_n, tie_up, out1);
}

GPSIQ_HD void spec_scan_epoch(double x, double d, const StepInfo& tab, int N, int T, int variant, double* ck,
                              size_t ck_stride, CarrSpec& out) {
    spec_scan_range(x, d, tab, N, T, 0, (N + T - 1) / T, variant, ck, ck_stride, out);
}

// Match an exact post-wrap state x (at range-relative sample n) against the two
// parity variants of a speculative run; on success returns the variant and the
// translation.  Shared by the chunk stitcher and the epoch chain.
GPSIQ_HD bool spec_match(double x, int n, double d, const CarrSpec& s0, const CarrSpec& s1, int& v, double& diff) {
    if (!(s0.margin > 0.0) || n != s0.n1) return false;
    const CarrSpec* s = &s0;
    v = 0;
    diff = x - s0.xw1;  // exact: both on the 2^-53 grid and close
    double q = diff * 0x1p52;
    if (d < 0.0 && q != (double) (long long) q && s1.margin > 0.0 && n == s1.n1) {
        s = &s1; v = 1;
        diff = x - s1.xw1;
        q = diff * 0x1p52;
    }
    const double ad = diff < 0.0 ? -diff : diff;
    return q == (double) (long long) q && ad < s->margin - 0x1p-50;
}

// ---- two-level speculation: chunks of G tiles ---------------------------------
// One speculative chain per whole epoch is still ~2000 serial segments.  So the
// epoch is cut into chunks of G tiles, each speculated independently from its
// own closed-form start estimate (spec_scan_range, all chunks of all epochs in
// parallel), and stitch_epoch then builds ONE epoch-level speculative trajectory
// P out of them with the same translate/verify step, chunk by chunk: it only
// runs the exact scan over the head of each chunk (up to its first wrap).  P is
// "the exact recurrence started from the epoch's estimated start phase"; the
// epoch chain (chain_epoch) treats it exactly like a directly scanned epoch.
struct ChunkInfo {  // per (epoch, channel, epoch-level variant, chunk)
    double delta;   // translation of the chunk's speculative run onto P
    int n1;         // tiles of the chunk starting before this epoch sample read the P plane; others: chunk plane + delta
    int variant;    // which chunk-level parity plane
};

// a0: estimated epoch start phase (the chunk-0 runs started from exactly this value).
// cs: chunk results [J][2]; ckP: this trajectory's P plane for head/fallback tiles.
GPSIQ_HD void stitch_epoch(double a0, double d, const StepInfo& tab, int N, int T, int G, int V,
                           const CarrSpec* cs, double* ckP, size_t ck_stride, ChunkInfo* ci, CarrSpec& outE) {
    const int ntiles = (N + T - 1) / T;
    const int J = (ntiles + G - 1) / G;
    double s = a0, marginE = 1.0;
    bool seen = false, usable = carr_step_speculable(d);
    outE.n1 = N;
    outE.xw1 = 0.0;
    for (int j = 0; j < J; j++) {
        const int t0 = j * G, t1 = (t0 + G < ntiles) ? t0 + G : ntiles;
        const int nstart = t0 * T;
        const int m_total = ((t1 * T < N) ? t1 * T : N) - nstart;
        ChunkInfo c;
        c.delta = 0.0; c.n1 = nstart + m_total; c.variant = 0;
        if (j == 0) {
            // the chunk-0 runs started from a0 itself: variant V *is* P (V flips the parity at the first wrap)
            const bool wrapped0 = cs[0].n1 < m_total;
            const CarrSpec& sp = cs[(wrapped0 && V == 1) ? 1 : 0];
            c.n1 = 0;
            c.variant = (wrapped0 && V == 1) ? 1 : 0;
            if (wrapped0) {
                seen = true;
                outE.n1 = sp.n1;
                outE.xw1 = sp.xw1;
                if (!(sp.margin > 0.0)) usable = false; else if (sp.margin < marginE) marginE = sp.margin;
            }
            s = sp.xend;
            ci[0] = c;
            continue;
        }
        // exact-in-P head of chunk j: up to and including its first wrap
        int n = 0, t = t0, remaining = 0;
        bool wrapped = false;
        for (; t < t1 && !wrapped; t++) {
            ckP[(size_t) t * ck_stride] = s;
            remaining = (T < N - t * T) ? T : N - t * T;
            while (remaining > 0 && !wrapped) {
                int steps;
                if (seen) steps = carr_advance<true>(s, d, tab, remaining, true, wrapped, marginE);
                else { double dummy = 1.0; steps = carr_advance<false>(s, d, tab, remaining, true, wrapped, dummy); }
                remaining -= steps;
                n += steps;
            }
        }
        if (wrapped) {
            if (!seen) {  // the epoch's first wrap happens in this head
                seen = true;
                if (V == 1) s = (s + 0x1p-53 < 1.0) ? s + 0x1p-53 : s - 0x1p-53;
                if (!(s >= 0.0 && s < 1.0)) usable = false;
                outE.n1 = nstart + n;
                outE.xw1 = s;
            }
            int v;
            double diff;
            if (spec_match(s, n, d, cs[j * 2], cs[j * 2 + 1], v, diff)) {
                const CarrSpec& sp = cs[j * 2 + v];
                c.delta = diff; c.n1 = nstart + n; c.variant = v;
                const double ad = diff < 0.0 ? -diff : diff;
                if (sp.margin - ad < marginE) marginE = sp.margin - ad;
                s = add_rn(sp.xend, diff);
                ci[j] = c;
                continue;
            }
            // the chunk's speculation does not fit: finish the chunk with the exact scan
            bool w;
            while (remaining > 0) remaining -= carr_advance<true>(s, d, tab, remaining, false, w, marginE);
            for (; t < t1; t++) {
                ckP[(size_t) t * ck_stride] = s;
                remaining = (T < N - t * T) ? T : N - t * T;
                while (remaining > 0) remaining -= carr_advance<true>(s, d, tab, remaining, false, w, marginE);
            }
        }
        ci[j] = c;  // every tile of the chunk reads the P plane
    }
    outE.xend = s;
    outE.margin = (seen && usable) ? marginE : -1.0;
    outE.pad = 0;
}


struct CarrInfo {   // per (epoch, channel): how the renderer obtains tile-start phases
    double delta;   // translation for tiles starting at or after n1
    int n1;         // tiles starting before n1 read the exact plane 0; N = all exact
    int variant;    // speculative plane (0/1) for the translated tiles
};

// ---- third level: groups of epochs ---------------------------------------------
// After stitching, the only serial work is one head scan per epoch (chain_epoch).  For
// long batches and for time-sliced multi-GPU runs even that is too much on the critical
// path, so the same step is applied once more: the epochs of a GROUP are chained from an
// ESTIMATED group start phase (all groups in parallel, group_chain_epoch with SPEC), which
// yields a group-level speculative trajectory with its own first-wrap state and margin;
// the final, truly serial chain then needs one head scan per GROUP.
struct GroupTrack {   // running state of a group-level speculative chain
    double margin;    // decision margin since the chain's first wrap
    double xw1;       // state right after that wrap
    int pos;          // its position: (epoch index within the group) * N + sample; -1: none yet
    int usable;
    TieEvent tie;     // first positive tie-wrap after pos (same units)
};

struct GroupInfo {    // per (group, channel): result of the final chain
    double delta;     // translation of the group trajectory onto the exact one
    double delta2;    // the same from tie_pos on (TieEvent)
    int pos;          // tiles at or after this (epoch-in-group * N + sample) are translated; earlier ones
                      // read the exact plane.  0x7fffffff: the whole group was chained exactly (fallback)
    int variant;
    int tie_pos;      // 0x7fffffff: none
    int pad;
};

// Scan from x over the tiles of one epoch starting at tile t (remaining = samples left of a partially
// scanned tile, or 0), writing the state at every tile start passed, until the first wrap or the epoch's
// end.  TRACK: margin-track every decision.
template <bool TRACK>
GPSIQ_HD void scan_epoch_head(double& x, double d, const StepInfo& tab, int N, int T, double* ck, size_t ck_stride,
                              int& t, int& n, int& remaining, bool& wrapped, bool stop_at_wrap, double& margin,
                              TieEvent* pt = nullptr, int pos_base = 0) {
    const int ntiles = (N + T - 1) / T;
    wrapped = false;
    for (;;) {
        while (remaining > 0 && !(wrapped && stop_at_wrap)) {
            bool w;
            const int steps = carr_advance<TRACK>(x, d, tab, remaining, stop_at_wrap, w, margin, pt, pos_base + n);
            remaining -= steps;
            n += steps;
            if (w) wrapped = true;
        }
        if ((wrapped && stop_at_wrap) || t >= ntiles) return;
        ck[(size_t) t * ck_stride] = x;
        remaining = (T < N - t * T) ? T : N - t * T;
        t++;
    }
}

// One epoch of a chain.  SPEC = false: the exact chain (this is chain_epoch).  SPEC = true: the chain is
// itself speculative (started from an estimate): once it has wrapped, every decision is margin-tracked
// into g, each epoch's own translation costs margin, and variant V flips the parity at the chain's
// first wrap.  eg = index of the epoch within its group.
template <bool SPEC>
GPSIQ_HD double group_chain_epoch(double x, double d, const StepInfo& tab, int N, int T, const CarrSpec& s0,
                                  const CarrSpec& s1, double* ck0, size_t ck_stride, CarrInfo& info, int& fell_back,
                                  GroupTrack* g, int V, int eg) {
    int n = 0, t = 0, remaining = 0;
    bool wrapped = false;
    double dummy = 1.0;
    if (SPEC && g->pos >= 0) scan_epoch_head<true>(x, d, tab, N, T, ck0, ck_stride, t, n, remaining, wrapped, true, g->margin, &g->tie, eg * N);
    else scan_epoch_head<false>(x, d, tab, N, T, ck0, ck_stride, t, n, remaining, wrapped, true, dummy);
    info.delta = 0.0;
    info.n1 = N;
    info.variant = 0;
    if (!wrapped) return x;  // the whole epoch was the head (low Doppler)
    if (SPEC && g->pos < 0) {  // the chain's first wrap
        if (V == 1) x = (x + 0x1p-53 < 1.0) ? x + 0x1p-53 : x - 0x1p-53;
        if (!(x >= 0.0 && x < 1.0)) g->usable = 0;
        g->pos = eg * N + n;
        g->xw1 = x;
    }
    {
        int v;
        double diff;
        if (spec_match(x, n, d, s0, s1, v, diff)) {
            info.delta = diff;
            info.n1 = n;
            info.variant = v;
            if (SPEC) {
                const double m = (v ? s1.margin : s0.margin) - (diff < 0.0 ? -diff : diff);
                if (m < g->margin) g->margin = m;
            }
            return add_rn(v ? s1.xend : s0.xend, diff);
        }
    }
    // the epoch's speculation does not fit: finish the epoch with the exact scan
    fell_back++;
    if (SPEC) scan_epoch_head<true>(x, d, tab, N, T, ck0, ck_stride, t, n, remaining, wrapped, false, g->margin, &g->tie, eg * N);
    else scan_epoch_head<false>(x, d, tab, N, T, ck0, ck_stride, t, n, remaining, wrapped, false, dummy);
    return x;
}

// Per-epoch inputs of the group-level chains, staged by the caller (shared memory on the device).
struct GroupEpoch {
    double d, phase0;
    CarrSpec s0, s1;   // the epoch's stitched (epoch-level) speculation results, both variants
    int active, reset;
};

// Level 3: speculative chain of one group of `count` epochs from x, an ESTIMATE of the phase at the
// group's first sample.  Head tiles go to ckHead (plane 4+V), per-epoch results to info/trace.
// outG.n1 holds the first-wrap position as (epoch-in-group * N + sample), count*N if the chain never wraps.
GPSIQ_HD void group_chain(double x, const GroupEpoch* ge, int count, int N, int T, int V, double* ckHead,
                          size_t tile_stride, size_t epoch_stride, CarrInfo* info, size_t info_stride, double* trace,
                          size_t trace_stride, CarrSpec& outG, TieEvent& tieG, int& fb) {
    GroupTrack g;
    g.margin = 1.0; g.xw1 = 0.0; g.pos = -1; g.usable = 1; g.tie.pos = -1; g.tie.k = 0;
    for (int eg = 0; eg < count; eg++) {
        CarrInfo inf;
        inf.delta = 0.0; inf.n1 = N; inf.variant = 0;
        if (ge[eg].active) {
            if (ge[eg].reset) { g.usable = 0; x = ge[eg].phase0; }       // re-seeded inside the group: not translatable
            if (!carr_step_translatable(ge[eg].d)) g.usable = 0;             // (a tie-capable step is scanned serially
            x = group_chain_epoch<true>(x, ge[eg].d, step_info(ge[eg].d), N, T, ge[eg].s0, ge[eg].s1,
                                        ckHead + (size_t) eg * epoch_stride, tile_stride, inf, fb, &g, V, eg);
        }
        info[(size_t) eg * info_stride] = inf;
        trace[(size_t) eg * trace_stride] = x;
    }
    outG.xw1 = g.xw1;
    outG.xend = x;
    outG.margin = (g.pos >= 0 && g.usable) ? g.margin : -1.0;
    outG.n1 = g.pos < 0 ? count * N : g.pos;
    outG.pad = 0;
    tieG = g.tie;
}

// Level 4: the final, exact chain over one group from the exact phase x: one head scan up to the group's
// first wrap, then the group trajectory translated -- or, if that does not fit, the rest chained exactly.
// Exact head / fallback tiles go to ckX (plane 6).  Returns the exact phase after the group.
GPSIQ_HD double group_final(double x, const GroupEpoch* ge, int count, int N, int T, const CarrSpec& sG0,
                            const CarrSpec& sG1, double* ckX, size_t tile_stride, size_t epoch_stride, CarrInfo* infoX,
                            size_t info_stride, const double* traceG0, const double* traceG1, double* trace,
                            size_t trace_stride, GroupInfo& gi, int& fb, const TieEvent& tG0, const TieEvent& tG1) {
    double dummy = 1.0;
    CarrInfo all_exact;
    all_exact.delta = 0.0; all_exact.n1 = N; all_exact.variant = 0;
    for (int eg = 0; eg < count; eg++) {
        infoX[(size_t) eg * info_stride] = all_exact;
        if (!ge[eg].active) { trace[(size_t) eg * trace_stride] = x; continue; }
        if (ge[eg].reset) x = ge[eg].phase0;
        int t = 0, n = 0, remaining = 0;
        bool wrapped = false;
        double* ck = ckX + (size_t) eg * epoch_stride;
        const double x_epoch = x;  // exact phase at the first sample of this epoch
        scan_epoch_head<false>(x, ge[eg].d, step_info(ge[eg].d), N, T, ck, tile_stride, t, n, remaining, wrapped, true, dummy);
        if (!wrapped) { trace[(size_t) eg * trace_stride] = x; continue; }
        const int pos = eg * N + n;
        int v;
        double diff;
        if (spec_match(x, pos, ge[eg].d, sG0, sG1, v, diff)) {
            const TieEvent& tv = v ? tG1 : tG0;
            const double diff2 = tie_shift(diff, tv);
            gi.delta = diff; gi.delta2 = diff2; gi.pos = pos; gi.variant = v; gi.pad = 0;
            gi.tie_pos = tv.pos < 0 ? 0x7fffffff : tv.pos;
            const double* tg = v ? traceG1 : traceG0;
            // the phase after epoch e2 follows the tie-wrap iff the event lies at or before that epoch's last step
            for (int e2 = eg; e2 < count; e2++)
                trace[(size_t) e2 * trace_stride] = add_rn(tg[(size_t) e2 * trace_stride], (gi.tie_pos <= (e2 + 1) * N) ? diff2 : diff);
            return add_rn(v ? sG1.xend : sG0.xend, diff2);
        }
        // The group's speculation does not fit: chain this and the remaining epochs exactly, each through its
        // own epoch-level speculation (one head scan + translation per epoch; an epoch scans serially only if
        // that does not fit either).
        fb++;
        x = x_epoch;
        for (int e2 = eg; e2 < count; e2++) {
            CarrInfo inf = all_exact;
            if (ge[e2].active) {
                if (e2 > eg && ge[e2].reset) x = ge[e2].phase0;
                x = group_chain_epoch<false>(x, ge[e2].d, step_info(ge[e2].d), N, T, ge[e2].s0, ge[e2].s1,
                                             ckX + (size_t) e2 * epoch_stride, tile_stride, inf, fb, nullptr, 0, 0);
            }
            infoX[(size_t) e2 * info_stride] = inf;
            trace[(size_t) e2 * trace_stride] = x;
        }
        gi.delta = 0.0; gi.delta2 = 0.0; gi.pos = 0x7fffffff; gi.variant = 0; gi.tie_pos = 0x7fffffff; gi.pad = 0;
        return x;
    }
    gi.delta = 0.0; gi.delta2 = 0.0; gi.pos = count * N; gi.variant = 0; gi.tie_pos = 0x7fffffff; gi.pad = 0;  // the group never wraps: every tile is in the exact plane
    return x;
}

// ---- fifth level: the whole batch ("slice") ---------------------------------------
// In a time-sliced multi-GPU run the exact chain of a slice sits on the ring: the next GPU cannot chain before this
// one has.  Level 4 costs one head scan per GROUP there (16 for a 1024-epoch slice).  The same step once more makes it
// ONE head scan per slice: the groups are chained from the ESTIMATED slice start (slice_chain_group, in the speculation
// phase, off the ring) into a slice-level speculative trajectory S' with its own first-wrap state and margin, recording
// the phase at which S' enters every group.  On the ring, slice_verify scans from the exact start phase up to the
// slice's first wrap only, matches it against S' and returns the exact end phase as xend' + D; the exact phase at the
// start of every group is then start'_g + D -- all groups at once -- and level 4 (group_final) is run for all groups in
// PARALLEL from those, after the hand-off, producing exactly what the serial chain would have produced.
// A slice whose trajectory does not wrap inside its first group, re-seeds a slot, or has a group whose speculation does
// not fit is not translatable (margin <= 0): it is chained serially, group by group, as before.

// Part of one group of the speculative slice-level chain: epochs ge[0..count) = the group's epochs eg0, eg0+1, ...
// x: phase of S' at the first sample of ge[0].  g: running state (GroupTrack.pos = slice-relative position of the
// slice's first wrap, epoch * N + sample; -1: none yet; GroupTrack.tie.pos = index of the GROUP holding S' first
// tie-wrap after that).  epoch0: index of the group's first epoch within the slice, gidx: index of the group.
// sG0 / sG1, tG0 / tG1: the group's own speculation (level 3), both variants.
// -> 1: the group's speculation was matched, x = phase of S' after the group's LAST sample;
//    0: none of these epochs wrapped (all head), x = phase after ge[count-1];
//   -1: the chain is not translatable from here on.
GPSIQ_HD int slice_chain_group(double& x, const GroupEpoch* ge, int count, int N, const CarrSpec& sG0, const CarrSpec& sG1,
                               const TieEvent& tG0, const TieEvent& tG1, GroupTrack& g, int V, int epoch0, int gidx, int eg0) {
    for (int k = 0; k < count; k++) {
        if (!ge[k].active) continue;
        if (ge[k].reset) return -1;                                  // re-seeded inside the slice: not translatable
        if (!carr_step_translatable(ge[k].d)) return -1;
        const StepInfo si = step_info(ge[k].d);
        bool wrapped = false;
        double dummy = 1.0;
        TieEvent th;
        th.pos = -1; th.k = 0;
        int n = 0;
        while (n < N && !wrapped)
            n += (g.pos >= 0) ? carr_advance<true>(x, ge[k].d, si, N - n, true, wrapped, g.margin, &th, 0)
                              : carr_advance<false>(x, ge[k].d, si, N - n, true, wrapped, dummy);
        if (th.pos >= 0 && g.tie.pos < 0) { g.tie.pos = gidx; g.tie.k = th.k; }   // (the wrap ending a head scan can tie)
        if (!wrapped) continue;                                      // the whole epoch was part of the head
        const int pos = (eg0 + k) * N + n;                           // group-relative, as group_chain records it
        if (g.pos < 0) {                                             // the slice's first wrap
            if (epoch0 != 0) return -1;                              // (it must lie inside the first group: slice_verify)
            if (V == 1) x = (x + 0x1p-53 < 1.0) ? x + 0x1p-53 : x - 0x1p-53;
            if (!(x >= 0.0 && x < 1.0)) return -1;
            g.pos = pos;
            g.xw1 = x;
        }
        int v;
        double diff;
        if (!spec_match(x, pos, ge[k].d, sG0, sG1, v, diff)) return -1;
        const double m = (v ? sG1.margin : sG0.margin) - (diff < 0.0 ? -diff : diff);
        if (m < g.margin) g.margin = m;
        // S' = G' + diff up to G's first tie-wrap, G' + diff2 after it; S' itself rounds that tie like G' if diff is
        // an even multiple of 2^-52 and the other way if it is odd
        const TieEvent& tv = v ? tG1 : tG0;
        const double diff2 = tie_shift(diff, tv);
        if (tv.pos >= 0 && g.tie.pos < 0) { g.tie.pos = gidx; g.tie.k = (diff2 != diff) ? -tv.k : tv.k; }
        x = add_rn(v ? sG1.xend : sG0.xend, diff2);
        return 1;
    }
    return 0;
}

// Exact head scan of a slice from its exact start phase over epochs ge[0..count) = epochs eg0, eg0+1, ... of its FIRST
// group, matched against the slice-level speculation (sS0 / sS1, tS0 / tS1: both variants).
// -> 1: matched; v = variant, diff / diff2 = translation before / after the trajectory's tie event (TieEvent.pos = a
//       GROUP index: the groups after it start from diff2), x = the exact phase after the slice's LAST sample;
//    0: none of these epochs wrapped, x = the exact phase after ge[count-1];
//   -1: no match (x is then somewhere inside the head: the caller restarts from the slice's start phase).
GPSIQ_HD int slice_verify(double& x, const GroupEpoch* ge, int count, int N, const CarrSpec& sS0, const CarrSpec& sS1,
                          const TieEvent& tS0, const TieEvent& tS1, int& v, double& diff, double& diff2, int eg0) {
    if (!(sS0.margin > 0.0)) return -1;
    double dummy = 1.0;
    for (int k = 0; k < count; k++) {
        if (!ge[k].active) continue;
        if (ge[k].reset) return -1;
        const StepInfo si = step_info(ge[k].d);
        bool wrapped = false;
        int n = 0;
        while (n < N && !wrapped) n += carr_advance<false>(x, ge[k].d, si, N - n, true, wrapped, dummy);
        if (!wrapped) continue;
        if (!spec_match(x, (eg0 + k) * N + n, ge[k].d, sS0, sS1, v, diff)) return -1;
        diff2 = tie_shift(diff, v ? tS1 : tS0);
        x = add_rn(v ? sS1.xend : sS0.xend, diff2);
        return 1;
    }
    return 0;
}

// Exact carrier phase at the start of tile t of epoch e (eg = its index within its group), composing the
// four levels: final chain (gi) -> group chain (inf) -> stitched trajectory P -> chunk runs.
// ck points at this (epoch, channel)'s entry of plane 0; `plane` elements per plane, `stride` elements
// between consecutive tiles.  Planes: 0,1 chunk runs; 2,3 P; 4,5 group-chain heads (variants); 6 exact.
// infG[0], infG[1]: the epoch's CarrInfo from the two group-chain variants; infX: from the exact fallback chain.
GPSIQ_HD double carr_tile_phase(const double* ck, size_t plane, size_t stride, int t, int T, int N, int G, int J, int eg,
                                const GroupInfo& gi, const CarrInfo& infG0, const CarrInfo& infG1, const CarrInfo& infX,
                                const ChunkInfo* ci, const CarrSpec* cs) {
    const int n0 = t * T;
    const size_t o = (size_t) t * stride;
    const bool all_exact = gi.pos == 0x7fffffff;
    if (!all_exact && eg * N + n0 < gi.pos) return ck[6 * plane + o];  // before the group's first wrap
    const CarrInfo& inf = all_exact ? infX : (gi.variant ? infG1 : infG0);
    const size_t headp = all_exact ? 6 : (size_t) (4 + gi.variant);
    double p;
    if (n0 < inf.n1 || inf.n1 >= N) {
        p = ck[headp * plane + o];
    } else {
        const int V = inf.variant;
        const ChunkInfo c = ci[V * J + t / G];
        if (n0 < c.n1) {
            p = ck[(size_t) (2 + V) * plane + o];
        } else {
            const int j = t / G;
            double q;
            if (c.variant == 1 && cs[j * 2 + 1].pad >= 0)   // derived variant: plane 0 shifted (cs = the chunk results [J][2])
                q = add_rn(ck[o], (double) derived_shift(cs[j * 2 + 1], n0 - j * G * T) * 0x1p-53);
            else
                q = ck[(size_t) c.variant * plane + o];
            p = add_rn(q, c.delta);
        }
        p = add_rn(p, inf.delta);
    }
    return all_exact ? p : add_rn(p, (eg * N + n0 >= gi.tie_pos) ? gi.delta2 : gi.delta);
}

// Everything a renderer needs to compose exact tile-start carrier phases (device pointers, one batch).
struct CarrLookup {
    const double* ck;        // [7][E][ntiles][C]
    size_t plane;            // elements per plane
    const GroupInfo* ginfo;  // [groups][C]
    const CarrInfo* infoG;   // [3][Ecap][C]: group-chain variants 0, 1 and the exact fallback chain
    size_t info_plane;       // Ecap * C
    const ChunkInfo* cinfo;  // [E][C][2][J]
    const CarrSpec* spec;    // [E][C][J][2] chunk results (the derived variant-1 planes are evaluated from them)
    int G, J, GP;            // chunk length (tiles), chunks per epoch, epochs per group
};

GPSIQ_HD double carr_lookup(const CarrLookup& L, int e, int c, int t, int T, int N, int C, int ntiles) {
    const size_t ec = (size_t) e * C + c;
    return carr_tile_phase(L.ck + (size_t) e * ntiles * C + c, L.plane, (size_t) C, t, T, N, L.G, L.J, e % L.GP,
                           L.ginfo[(size_t) (e / L.GP) * C + c], L.infoG[ec], L.infoG[L.info_plane + ec],
                           L.infoG[2 * L.info_plane + ec], L.cinfo + ec * 2 * L.J, L.spec + ec * 2 * L.J);
}

// (3) exact chaining of one epoch from the exact start x; returns the exact end state.
GPSIQ_HD double chain_epoch(double x, double d, const StepInfo& tab, int N, int T, const CarrSpec& s0,
                            const CarrSpec& s1, double* ck0, size_t ck_stride, CarrInfo& info, int& fell_back) {
    return group_chain_epoch<false>(x, d, tab, N, T, s0, s1, ck0, ck_stride, info, fell_back, nullptr, 0, 0);
}

}  // namespace gpsiq
