// synth_line.cuh — k_synth_line: the production synthesis kernel (plutogpssim.c:2689-2756),
// with k_line_anchor (tile anchors + safety check) before and k_line_patch / k_line_apply after it.
//
// Per channel and sample the main kernel does exactly: two 64-bit adds (the two NCO lines, kept in
// "split-word" form: the HIGH 32-bit word of each accumulator is the shared-memory address material,
// the low word the fraction), one AND/OR (carrier table index, plutogpssim.c:2697 -> LUT address; the
// chip index of plutogpssim.c:2737 needs none: the high word IS the byte address), two shared-memory
// loads (amplitude LUT entry = (int)(cos|sin * gain) of plutogpssim.c:2701-2702 packed as Q<<16 + I;
// chip/NAV sign as a +-1 byte) and one multiply-add into the packed accumulator
// (plutogpssim.c:2705-2706): 8 instructions.  No branch, no per-run state, no correction array: see
// line_check.cuh for why a straight line per 1024-sample tile is exact once k_line_anchor has cleared
// the tile; the split-word form truncates the line by < LN_KF / LN_KG units, which the check covers.
//
//   tile        1024 samples; anchor = exact NCO states at its first sample (from the scan phases)
//   warp-block  512 samples: lane l renders samples l, l+32, ..., l+480 of the block, so the 32 lanes
//               of one shared-memory request touch <= 32 consecutive LUT entries (no bank conflict)
//               and <= 16 consecutive chip bytes (broadcast), and every store instruction of a warp
//               writes 128 contiguous bytes
//   CTA         512 threads, persistent: a contiguous range of the batch's tiles; tables of the
//               current epoch resident in shared memory: per slot the 512-entry amplitude LUT and the
//               chip-sign table in 4 variants (NAV polarity before / after the one code-period wrap a
//               tile can contain, plutogpssim.c:2710-2733), extended past chip 1022 so that the code
//               line never has to wrap inside a tile
//   chip index  the high bits of G hold the shared-memory byte address of the slot's table + variant
//               offset + chip, so the load needs no address arithmetic
// More than LN_CG slots: the CTA makes one pass per channel group over its range and keeps the raw
// packed sums in the output buffer between passes.
#pragma once

#include "line_check.cuh"

// Geometry (measured on B200, profiles/r02_r_geometry_sweep.txt): a lane renders LN_RUN = 32 samples of a slot before it
// moves to the next slot, so a warp-block is a whole 1024-sample tile and the ~20 instructions of per-slot set-up are
// paid once per 32 samples instead of once per 16 (round 1: LN_RUN 16, 512 threads, 48 registers: 1.35 ms per 1024-epoch
// launch; now 1.14 ms).  The 32 accumulators need 64 registers: 448 threads per CTA, two CTAs per SM = 7/8 of the
// register file, the rest is left to the scan kernels that run beside it.  (-D overrides for experiments.)
#ifndef LN_RUN
#define LN_RUN 32
#endif
#define LN_WB (32 * LN_RUN)       // samples per warp-block
#define LN_TILE 1024
#ifndef LN_THREADS
#define LN_THREADS 448
#endif
#define LN_WARPS (LN_THREADS / 32)
#define LN_CG 12                  // slots resident at a time
#define LN_VS 1536                // chip-sign entries per variant: 1023 + 512 (code_step <= 0.5) + 1
#define LN_CHUNK 32               // tiles per hazard chunk (= lanes)
// Split-word lines: a lane's LN_RUN samples of a warp-block are evaluated as
//     X_j = (F0 >> LN_XSH) + j * DX,  DX = (32 dF) >> LN_XSH (arithmetic)   carrier: index = bits [10:2] of X's high word
//     Y_j = (G0 >> LN_YSH) + j * DY,  DY = (32 dG) >> LN_YSH                code: byte address = Y's high word
// from the lane's exact 64-bit start F0 = anchor + m0 dF.  Both truncations round down, so the value the
// kernel floors lies in (line - LN_K*, line]: k_line_anchor widens the lower side of its hazard window by LN_K*.
#define LN_XSH 21
#define LN_YSH 15
#define LN_KF ((int64_t) LN_RUN << LN_XSH)
#define LN_KG ((int64_t) LN_RUN << LN_YSH)

#define LN_DBG_FORCE_CHUNK 1      // cfg.reserved[1] bits (tests): treat every chunk as flagged,
#define LN_DBG_FORCE_TILE 2       //   every tile as a hazard (all samples re-checked by k_line_patch),
#define LN_DBG_PERTURB 4          //   and shift the anchors the main kernel uses (patches must repair it)

namespace gpsiq {

struct LinePatch { uint32_t sample; int32_t delta; };  // batch-relative sample index, packed (right - wrong)

__host__ __device__ inline int ln_groups(int C) { return (C + LN_CG - 1) / LN_CG; }
__host__ __device__ inline int ln_group_slots(int C) { const int g = ln_groups(C); return (C + g - 1) / g; }
__host__ __device__ inline size_t ln_smem_bytes(int C) {
    const int CG = ln_group_slots(C);
    return (size_t) CG * 4 * LN_VS + 2048 /* LUT alignment */ + (size_t) CG * 2048 + 16 * 32 /* steps */ +
           LN_WARPS * 16 * 16 /* anchors */ + 128;
}

// what k_synth_line evaluates for tile-relative sample n of a tile anchored at (aF, aG): carrier table index and
// chip-table index (k_line_patch compares exactly this with the literal recurrence)
GPSIQ_HD uint64_t ln_split_dx(uint64_t dF) { return (uint64_t) ((int64_t) dF >> (LN_XSH - 5)); }
GPSIQ_HD uint64_t ln_split_dy(uint64_t dG) { return dG >> (LN_YSH - 5); }
GPSIQ_HD void ln_kernel_index(uint64_t aF, uint64_t aG, uint64_t dF, uint64_t dG, uint32_t n, uint32_t& ci, uint32_t& gi) {
    const uint32_t m0 = (n & ~(uint32_t) (LN_WB - 1)) | (n & 31u);  // the lane's first sample of the warp-block
    const uint32_t j = (n >> 5) & (uint32_t) (LN_RUN - 1);
    const uint64_t X = ((aF + (uint64_t) m0 * dF) >> LN_XSH) + (uint64_t) j * ln_split_dx(dF);
    const uint64_t Y = ((aG + (uint64_t) m0 * dG) >> LN_YSH) + (uint64_t) j * ln_split_dy(dG);
    ci = (uint32_t) (X >> (LN_FBITS - LN_XSH)) & 511u;
    gi = (uint32_t) (Y >> (LN_GBITS - LN_YSH));
}

// Integer carrier NCO (GPSIQ_CARRIER_INT32; the reference's #else branches, plutogpssim.c:2699, 2748): the phase is
// a uint32 with 2^25 counts per cycle (index = bits 24:16), advanced by an integer step per sample -- EXACTLY a
// line.  F = u << 39 puts the index into bits 63:55 like the float carrier's F, the counts above bit 24 fall off
// the top (the reference masks them, & 0x1ff), and the low 39 bits of F and dF are zero, so both split-word
// truncations are exact: this carrier needs no hazard test, no scan and no patch.
GPSIQ_HD uint64_t ln_int_fixed(uint32_t u) { return (uint64_t) u << 39; }
GPSIQ_HD uint64_t ln_carr_slope_mode(double carr_step, int int_carrier) {
    return int_carrier ? (uint64_t) (int64_t) (int32_t) carr_step << 39 : ln_carr_slope(carr_step);
}

// ---- anchors and the safety check -----------------------------------------------------
// k_line_anchor   one warp per (epoch, slot), lane = tile: the tile anchors
//                   anch[(e*ntiles + t)*C + c] = { F (carrier), G (code) + chip-table variant offset }
//                 and the spread [lo, hi] of the carrier anchors around ONE line over the whole epoch (from the
//                 epoch's first anchor).  The carrier anchors are the exact tile-start phases of the scan phases (or
//                 the integer carrier's closed form); the CODE anchors are closed form too: the code NCO restarts at
//                 every epoch from an exactly known phase (computeCodePhase, plutogpssim.c:1770), so after n steps
//                 the true phase lies within ln_eps(0, n) of the straight fixed-point line from code_phase0 -- no
//                 code-NCO scan is needed at all, the bound simply widens the hazard window along the epoch.
// k_line_check    one THREAD per (epoch, slot): both lines of the whole epoch against their windows (line_hazard,
//                 the Euclid-like descent: ~10^4 instructions, so packed 32 per warp and run once per epoch instead
//                 of once per 32-tile chunk).  The ~1 % of (epoch, slot) pairs it cannot clear go to a list.
// k_line_refine   one warp per listed pair: per chunk of 32 tiles (lane = chunk), then per tile of the flagged chunks
//                 (lane = tile, each from its own anchor); the (tile, slot) pairs that still cannot be cleared go to
//                 hazlist for k_line_patch.
struct LineTile { uint64_t FA, GA, Fs, Gs; };
struct LineEpoch { uint64_t F0; int64_t lo, hi; int active, pad; };  // per (epoch, slot), written by k_line_anchor

// Closed-form code anchor of tile-relative sample n0 (epoch sample index): the epoch's code line G0 + n0*dG in exact
// 128-bit integer arithmetic, reduced to (chips in [0,1023)) * 2^47 + the number of code-period wraps so far.
GPSIQ_HD void ln_code_line(uint64_t G0, uint64_t dG, uint32_t n0, uint64_t& GA, int& wraps) {
#if defined(__CUDA_ARCH__)
    const uint64_t plo = (uint64_t) n0 * dG, phi = __umul64hi((uint64_t) n0, dG);
#else
    const unsigned __int128 pr = (unsigned __int128) n0 * dG;
    const uint64_t plo = (uint64_t) pr, phi = (uint64_t) (pr >> 64);
#endif
    const uint64_t lo = plo + G0, hi = phi + (lo < plo ? 1u : 0u);
    const uint32_t chips = (uint32_t) ((hi << (64 - LN_GBITS)) | (lo >> LN_GBITS));  // < 2^30: n0 < 2^31, step <= 0.5
    const uint32_t w = chips / 1023u;
    wraps = (int) w;
    GA = ((uint64_t) (chips - w * 1023u) << LN_GBITS) | (lo & ((1ULL << LN_GBITS) - 1));
}

__device__ __forceinline__ LineTile ln_tile_anchor(const gpsiq_chan_desc& d, const CarrLookup& carr,
                                                   const uint32_t* __restrict__ ustart, uint64_t G0, uint64_t dG, int e, int c,
                                                   int t, int C, int N, int ntiles, int dbg) {
    LineTile r;
    if (ustart)  // integer carrier: closed form from the epoch's start phase
        r.FA = ln_int_fixed(ustart[(size_t) e * C + c] + (uint32_t) (int32_t) d.carr_step * (uint32_t) (t * LN_TILE));
    else
        r.FA = ln_carr_fixed(carr_lookup(carr, e, c, t, LN_TILE, N, C, ntiles));
    int wraps;
    ln_code_line(G0, dG, (uint32_t) t * LN_TILE, r.GA, wraps);
    // NAV polarity at the tile start and after the next code-period wrap (plutogpssim.c:2714-2733)
    const int wr = wraps + d.ms0 % 20;
    const uint32_t pol0 = (uint32_t) (d.navbits >> ((wr / 20) & 63)) & 1u;
    const uint32_t pol1 = (uint32_t) (d.navbits >> (((wr + 1) / 20) & 63)) & 1u;
    r.Fs = r.FA;
    r.Gs = r.GA + ((uint64_t) ((pol0 * 2 + pol1) * LN_VS) << LN_GBITS);
    if (dbg & LN_DBG_PERTURB) { r.Fs += (uint64_t) (1 + t % 3) << 49; r.Gs += (uint64_t) (t % 5) << 42; }
    return r;
}

__global__ void __launch_bounds__(128)
k_line_anchor(const gpsiq_chan_desc* __restrict__ desc, const CarrLookup carr, const uint32_t* __restrict__ ustart,
              const int* __restrict__ amp_sum, int* __restrict__ step_flag, ulonglong2* __restrict__ anch,
              LineEpoch* __restrict__ recs, uint32_t* __restrict__ hazlist, int* __restrict__ counters, int haz_cap,
              int E, int C, int N, int ntiles, int dbg) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int c = w % C, e = w / C;
    if (e >= E) return;
    const gpsiq_chan_desc d = desc[(size_t) e * C + c];
    LineEpoch rec;
    rec.F0 = 0; rec.lo = 0; rec.hi = 0; rec.active = 0; rec.pad = 0;
    if (d.prn <= 0 || amp_sum[e] > 32767 || step_flag[e]) {  // inactive slot / epoch rendered by k_synth_lanes
        for (int t = lane; t < ntiles; t += 32) anch[((size_t) e * ntiles + t) * C + c] = make_ulonglong2(0, 0);
        if (lane == 0) recs[(size_t) e * C + c] = rec;
        return;
    }
    const bool int_carrier = ustart != nullptr;
    const uint64_t dF = ln_carr_slope_mode(d.carr_step, int_carrier), dG = ln_code_slope(d.code_step);
    const uint64_t G0 = ln_code_fixed(d.code_phase0);
    uint64_t F0 = 0;
    int64_t lo = 0, hi = 0;
    for (int t0 = 0; t0 < ntiles; t0 += 32) {
        const int t = t0 + lane;
        LineTile a;
        a.FA = 0;
        if (t < ntiles) {
            a = ln_tile_anchor(d, carr, ustart, G0, dG, e, c, t, C, N, ntiles, dbg);
            anch[((size_t) e * ntiles + t) * C + c] = make_ulonglong2(a.Fs, a.Gs);
        }
        if (t0 == 0) F0 = __shfl_sync(0xffffffffu, a.FA, 0);
        if (t < ntiles) {
            const int64_t dl = (int64_t) (a.FA - (F0 + (uint64_t) t * LN_TILE * dF));
            lo = min(lo, dl); hi = max(hi, dl);   // (tile 0 contributes 0)
        }
    }
#pragma unroll
    for (int s = 16; s; s >>= 1) {
        lo = min(lo, (int64_t) __shfl_xor_sync(0xffffffffu, lo, s));
        hi = max(hi, (int64_t) __shfl_xor_sync(0xffffffffu, hi, s));
    }
    if (lane == 0) {
        rec.F0 = F0; rec.lo = lo; rec.hi = hi; rec.active = 1;
        recs[(size_t) e * C + c] = rec;
    }
    if (dbg & LN_DBG_FORCE_TILE) {  // test hook: every tile goes through the literal-recurrence patch path
        for (int t = lane; t < ntiles; t += 32) {
            const int slot = atomicAdd(&counters[0], 1);
            if (slot < haz_cap) hazlist[slot] = (uint32_t) (e * ntiles + t) * 32u + (uint32_t) c;
            else atomicOr(&step_flag[e], 4);
        }
    }
}

// windows of the two lines over samples [n0, n0 + n) of the epoch, relative to the epoch's lines
__device__ __forceinline__ bool ln_range_hazard(const LineEpoch& r, uint64_t dF, uint64_t G0, uint64_t dG, bool int_carrier,
                                                uint32_t n0, uint32_t n) {
    const int64_t eF = ln_eps(1, LN_TILE);                  // inside a tile, from the tile's own exact carrier anchor
    const int64_t eG = ln_eps(0, (int64_t) n0 + n);         // the code line runs from the epoch's first sample
    if (!int_carrier && line_hazard(r.F0 + (uint64_t) n0 * dF, dF, LN_FBITS, n, r.lo - eF - LN_KF, r.hi + eF)) return true;
    return line_hazard(G0 + (uint64_t) n0 * dG, dG, LN_GBITS, n, -eG - LN_KG, eG);  // (only G mod 2^47 matters)
}

__global__ void __launch_bounds__(128)
k_line_check(const gpsiq_chan_desc* __restrict__ desc, const LineEpoch* __restrict__ recs, uint32_t* __restrict__ elist,
             int* __restrict__ counters, int EC, int N, int int_carrier, int dbg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= EC) return;
    const LineEpoch r = recs[i];
    if (!r.active || (dbg & LN_DBG_FORCE_TILE)) return;
    const gpsiq_chan_desc d = desc[i];
    const uint64_t dF = ln_carr_slope_mode(d.carr_step, int_carrier), dG = ln_code_slope(d.code_step);
    const bool hz = (dbg & LN_DBG_FORCE_CHUNK) || ln_range_hazard(r, dF, ln_code_fixed(d.code_phase0), dG, int_carrier, 0u, (uint32_t) N);
    if (hz) elist[atomicAdd(&counters[3], 1)] = (uint32_t) i;
}

__global__ void __launch_bounds__(128)
k_line_refine(const gpsiq_chan_desc* __restrict__ desc, const LineEpoch* __restrict__ recs, const CarrLookup carr,
              const uint32_t* __restrict__ ustart, const uint32_t* __restrict__ elist, int* __restrict__ step_flag,
              uint32_t* __restrict__ hazlist, int* __restrict__ counters, int haz_cap, int C, int N, int ntiles, int dbg) {
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * 4;
    const int count = counters[3];
    const bool int_carrier = ustart != nullptr;
    const int chunks = (ntiles + LN_CHUNK - 1) / LN_CHUNK;
    for (int it = blockIdx.x * 4 + (threadIdx.x >> 5); it < count; it += nwarps) {
        const int i = (int) elist[it];
        const int e = i / C, c = i - e * C;
        const gpsiq_chan_desc d = desc[i];
        const LineEpoch r = recs[i];
        const uint64_t dF = ln_carr_slope_mode(d.carr_step, int_carrier), dG = ln_code_slope(d.code_step);
        const uint64_t G0 = ln_code_fixed(d.code_phase0);
        if (lane == 0) atomicAdd(&counters[2], 1);  // diagnostic: (epoch, slot) pairs refined
        for (int ch0 = 0; ch0 < chunks; ch0 += 32) {
            const int ch = ch0 + lane;
            bool hz = false;
            if (ch < chunks) {
                const uint32_t n0 = (uint32_t) ch * LN_CHUNK * LN_TILE;
                const uint32_t n = (uint32_t) min(LN_CHUNK * LN_TILE, N - (int) n0);
                hz = (dbg & LN_DBG_FORCE_CHUNK) || ln_range_hazard(r, dF, G0, dG, int_carrier, n0, n);
            }
            uint32_t flagged = __ballot_sync(0xffffffffu, hz);
            // ---- tile-level check of the flagged chunks, each tile from its own anchor
            while (flagged) {
                const int jj = __ffs(flagged) - 1;
                flagged &= flagged - 1;
                const int t = (ch0 + jj) * LN_CHUNK + lane;
                if (t >= ntiles) continue;
                const LineTile a = ln_tile_anchor(d, carr, ustart, G0, dG, e, c, t, C, N, ntiles, 0);
                const int len = min(LN_TILE, N - t * LN_TILE);
                const int64_t eF = ln_eps(1, LN_TILE), eG = ln_eps(0, (int64_t) t * LN_TILE + len);
                const bool hzt = (!int_carrier && line_hazard(a.FA, dF, LN_FBITS, (uint64_t) len, -eF - LN_KF, eF)) ||
                                 line_hazard(a.GA, dG, LN_GBITS, (uint64_t) len, -eG - LN_KG, eG);
                if (hzt) {
                    const int slot = atomicAdd(&counters[0], 1);
                    if (slot < haz_cap) hazlist[slot] = (uint32_t) (e * ntiles + t) * 32u + (uint32_t) c;
                    else atomicOr(&step_flag[e], 4);  // list full: the epoch is re-rendered by k_synth_lanes
                }
            }
        }
    }
}

// ---- k_line_patch ------------------------------------------------------------------
// One thread per listed (tile, slot): the literal recurrences of plutogpssim.c:2697-2746 from the exact
// tile-start state, compared sample by sample with what k_synth_line computes from the anchors.
__global__ void __launch_bounds__(128)
k_line_patch(const gpsiq_chan_desc* __restrict__ desc, const int32_t* __restrict__ lutp, const CarrLookup carr,
             const uint32_t* __restrict__ ustart,
             const ulonglong2* __restrict__ anch, const int8_t* __restrict__ chips4,
             const uint32_t* __restrict__ hazlist, int* __restrict__ counters, int haz_cap,
             LinePatch* __restrict__ patches, int patch_cap, int* __restrict__ step_flag, int C, int N, int ntiles) {
    const int count = min(counters[0], haz_cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const uint32_t item = hazlist[i];
        const int c = item & 31;
        const int tg = item >> 5;
        const int e = tg / ntiles, t = tg - e * ntiles;
        const size_t ec = (size_t) e * C + c;
        const gpsiq_chan_desc d = desc[ec];
        const size_t o = ((size_t) e * ntiles + t) * C + c;
        const int len = min(LN_TILE, N - t * LN_TILE);
        // exact code-NCO state at the tile start: the recurrence restarts at every epoch (plutogpssim.c:1770), so it is
        // the exact fast-forward over the t * 1024 samples before the tile
        double cp = d.code_phase0;
        int wraps0 = 0;
        nco_advance<NCO_CODE>(cp, d.code_step, step_info(d.code_step), t * LN_TILE, wraps0);
        const bool int_carrier = ustart != nullptr;
        double ph = int_carrier ? 0.0 : carr_lookup(carr, e, c, t, LN_TILE, N, C, ntiles);
        const uint32_t ustep = (uint32_t) (int32_t) d.carr_step;
        uint32_t u = int_carrier ? ustart[ec] + ustep * (uint32_t) (t * LN_TILE) : 0u;
        const int wr = wraps0 + d.ms0 % 20;
        int kbit = wr / 20, icode = wr - kbit * 20;
        const ulonglong2 a = anch[o];
        const uint64_t dF = ln_carr_slope_mode(d.carr_step, int_carrier), dG = ln_code_slope(d.code_step);
        const int32_t* lut = lutp + ec * 512;
        const int8_t* chipv = chips4 + (size_t) d.prn * 4 * LN_VS;  // variants of this PRN
        const int8_t* chip0 = chips4 + (size_t) d.prn * 4 * LN_VS;  // variant 0 = polarity (0,0): +1 iff chip == 0
        for (int n = 0; n < len; n++) {
            // exact (plutogpssim.c:2697, 2701-2702, 2732, 2737)
            const int it = int_carrier ? (int) ((u >> 16) & 0x1ff)   // plutogpssim.c:2699
                                       : min(__double2int_rd(__dmul_rn(ph, 512.0)), 511);
            const int chipi = __double2int_rz(cp);
            const int chipbit = chip0[chipi] > 0 ? 0 : 1;
            const int nav = (int) (d.navbits >> (kbit & 63)) & 1;
            const int32_t right = (chipbit == nav) ? lut[it] : -lut[it];
            // what k_synth_line adds for this slot
            uint32_t ci, gi;
            ln_kernel_index(a.x, a.y, dF, dG, (uint32_t) n, ci, gi);
            const int32_t wrong = lut[ci] * (int32_t) chipv[gi];
            if (right != wrong) {
                const int slot = atomicAdd(&counters[1], 1);
                if (slot < patch_cap) {
                    LinePatch p; p.sample = (uint32_t) ((size_t) e * N + (size_t) t * LN_TILE + n); p.delta = right - wrong;
                    patches[slot] = p;
                } else {
                    atomicOr(&step_flag[e], 4);
                }
            }
            int wdummy = 0;
            if (nco_step<NCO_CODE>(cp, d.code_step, wdummy)) { if (++icode >= 20) { icode = 0; kbit++; } }
            if (int_carrier) u += ustep;                             // plutogpssim.c:2748
            else nco_step<NCO_CARRIER>(ph, d.carr_step, wdummy);
        }
    }
}

// packed sum (Q << 16) + I with signed I  <->  int16 pair as stored (plutogpssim.c:2754-2755)
__device__ __forceinline__ uint32_t ln_pack(uint32_t acc) { return acc + ((acc & 0x8000u) << 1); }
__device__ __forceinline__ uint32_t ln_unpack(uint32_t w) { return w - ((w & 0x8000u) << 1); }

__global__ void __launch_bounds__(128)
k_line_apply(const LinePatch* __restrict__ patches, const int* __restrict__ counters, int patch_cap,
             uint32_t* __restrict__ iq, unsigned long long s_lo, unsigned long long s_hi,
             unsigned long long* __restrict__ totals) {
    const int count = min(counters[1], patch_cap);
    if (totals && blockIdx.x == 0 && threadIdx.x < 3) totals[threadIdx.x] += (unsigned long long) counters[threadIdx.x];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const LinePatch p = patches[i];
        if (p.sample < s_lo || p.sample >= s_hi) continue;  // samples of this sub-batch only
        uint32_t* w = iq + p.sample;
        uint32_t old = *w, seen;
        do {
            seen = old;
            old = atomicCAS(w, seen, ln_pack(ln_unpack(seen) + (uint32_t) p.delta));
        } while (old != seen);
    }
}

// ---- k_synth_line --------------------------------------------------------------------
__device__ __forceinline__ int32_t ln_lds_s8(uint32_t saddr) {
    int32_t v;
    asm volatile("ld.shared.s8 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
// LUT entry of the 2 KB-aligned table at shared address base, index taken from bits [10:2] of the carrier
// accumulator's high word: one LOP3 for the address
__device__ __forceinline__ int32_t ln_lds_lut(uint32_t base, uint32_t xh) {
    int32_t v;
    uint32_t addr;
    asm volatile("lop3.b32 %0, %1, 0x7fc, %2, 0xea;" : "=r"(addr) : "r"(xh), "r"(base));  // (xh & 0x7fc) | base
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ ulonglong2 ln_lds_v2(uint32_t saddr) {
    ulonglong2 v;
    asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(saddr));
    return v;
}

// ---- TMA bulk copies (cp.async.bulk, global -> shared, completion on an mbarrier): the per-epoch tables are
// staged by a handful of copy instructions instead of ~340 load/store instructions per warp
__device__ __forceinline__ void ln_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void ln_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ln_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// -> false if the phase did not complete within a (very long) bounded spin: the caller flags an error
__device__ __forceinline__ bool ln_mbar_wait(uint32_t bar, uint32_t parity) {
    for (int spin = 0; spin < (1 << 22); spin++) {
        uint32_t ok;
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    return false;
}

// One CTA per unit of LN_UNIT consecutive tiles of one epoch.  The CTAs are deliberately NOT persistent:
// they retire every ~100 us, so the hardware block scheduler can place the small latency-bound kernels of
// the next batch's carrier chain (launched on higher-priority streams) beside this kernel as slots free up.
// A persistent variant with its own tile scheduler was ~3 % faster alone but starved those kernels.
#ifndef LN_UNIT
#define LN_UNIT 42   // tiles per CTA: three per warp
#endif
#ifndef LN_MAXNREG
#define LN_MAXNREG 64
#endif

// 48 registers: two CTAs use 3/4 of an SM's register file, so that the small latency-bound kernels of the
// next batch's carrier chain (stitch / group / final, code scan) can run beside it
__global__ void __maxnreg__(LN_MAXNREG)
k_synth_line(const gpsiq_chan_desc* __restrict__ desc, const int32_t* __restrict__ lutp,
             const int8_t* __restrict__ chips4, const ulonglong2* __restrict__ anch,
             const int* __restrict__ amp_sum, const int* __restrict__ step_flag, int16_t* __restrict__ iq,
             int E, int C, int N, int ntiles, int int_carrier, int* __restrict__ err) {
    extern __shared__ __align__(16) unsigned char ln_raw[];
    const int CG = ln_group_slots(C), ngroups = ln_groups(C);
    int8_t* s_chip = (int8_t*) ln_raw;                                   // [CG][4][LN_VS]
    const uint32_t chip_saddr = (uint32_t) __cvta_generic_to_shared(s_chip);
    // the LUTs start on a 2 KB boundary of the shared window, so that (index bits) | base is the entry address
    const uint32_t lut_saddr = (chip_saddr + (uint32_t) CG * 4 * LN_VS + 2047u) & ~2047u;
    int32_t* s_lut = (int32_t*) (ln_raw + (lut_saddr - chip_saddr));     // [CG][512]
    ulonglong2* s_step = (ulonglong2*) (s_lut + (size_t) CG * 512);      // [16][2] {dF, dG}, {DX, DY}; dG == 0: inactive
    ulonglong2* s_anch = s_step + 32;                                    // [LN_WARPS][16]
    int* s_prn = (int*) (s_anch + LN_WARPS * 16);                        // [16] PRN whose chip tables are resident
    uint64_t* s_bar = (uint64_t*) (s_prn + 16);                          // mbarrier of the table copies
    uint32_t* s_tx = (uint32_t*) (s_bar + 1);                            // bytes in flight for the current fill
    const uint32_t bar_saddr = (uint32_t) __cvta_generic_to_shared(s_bar);
    uint32_t bar_phase = 0;
    if (chip_saddr + (uint32_t) CG * 4 * LN_VS > (1u << (64 - LN_GBITS))) {  // the chip address must fit G's index field
        if (threadIdx.x == 0) atomicExch(err, 0x40000000);
        return;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    ulonglong2* my_anch = s_anch + warp * 16;
    const uint32_t step_saddr = (uint32_t) __cvta_generic_to_shared(s_step);
    const uint32_t anch_saddr = (uint32_t) __cvta_generic_to_shared(my_anch);
    const int wb_epoch = (N + LN_WB - 1) / LN_WB;
    const int upe = (ntiles + LN_UNIT - 1) / LN_UNIT;  // units per epoch
    const unsigned int nunits = (unsigned int) E * upe;
    constexpr int WPT = LN_TILE / LN_WB;               // warp-blocks per tile
    if (threadIdx.x < 16) s_prn[threadIdx.x] = -1;
    if (threadIdx.x == 0) ln_mbar_init(bar_saddr, 1);
    int cur_e = -1;
    (void) E;

    // a CTA takes a contiguous range of units (one unit each unless the launch caps the grid): consecutive units
    // of one epoch reuse the resident tables
    const unsigned int unit_lo = (unsigned int) ((unsigned long long) nunits * blockIdx.x / gridDim.x);
    const unsigned int unit_hi = (unsigned int) ((unsigned long long) nunits * (blockIdx.x + 1) / gridDim.x);
    for (unsigned int unit = unit_lo; unit < unit_hi; unit++) {
        __syncthreads();  // everyone is done with the previous unit's tables
        const int e = (int) (unit / upe);
        const int t0 = (int) (unit - (unsigned int) e * upe) * LN_UNIT;
        const int t1 = min(ntiles, t0 + LN_UNIT);
        if (amp_sum[e] > 32767 || step_flag[e]) continue;  // rendered by k_synth_lanes (uniform per CTA)
        const gpsiq_chan_desc* de = desc + (size_t) e * C;
        uint32_t* out_epoch = reinterpret_cast<uint32_t*>(iq) + (size_t) e * N;
        const int wb0 = t0 * WPT, wb1 = min(t1 * WPT, wb_epoch);
        for (int g = 0; g < ngroups; g++) {
            const int c0 = g * CG, nc = min(CG, C - c0);
            if (ngroups > 1 || e != cur_e) {
                if (g > 0) __syncthreads();  // everyone is done with the previous group's tables
                // Tables of the epoch by TMA bulk copies, one lane of warp 0 per slot: the amplitude LUT, and the chip
                // tables of the slots whose PRN differs from what is resident.  (The top-of-unit barrier ordered all
                // reads of the previous tables before these writes.)
                if (warp == 0) {
                    int prn = 0;
                    if (lane < nc) prn = de[c0 + lane].prn;
                    const bool need_lut = prn > 0;
                    const bool need_chip = prn > 0 && prn <= 32 && prn != s_prn[lane];
                    const uint32_t mc = __ballot_sync(0xffffffffu, need_chip), ml = __ballot_sync(0xffffffffu, need_lut);
                    const uint32_t total = (uint32_t) __popc(mc) * (4u * LN_VS) + (uint32_t) __popc(ml) * 2048u;
                    if (lane == 0) {
                        *s_tx = total;
                        if (total) ln_mbar_expect_tx(bar_saddr, total);
                    }
                    __syncwarp();
                    if (need_chip)
                        ln_bulk_g2s(chip_saddr + (uint32_t) lane * 4 * LN_VS, chips4 + (size_t) prn * 4 * LN_VS, 4 * LN_VS, bar_saddr);
                    if (need_lut)
                        ln_bulk_g2s(lut_saddr + (uint32_t) lane * 2048, lutp + ((size_t) e * C + c0 + lane) * 512, 2048, bar_saddr);
                }
                __syncthreads();  // all reads of s_prn above are done
                if (threadIdx.x < 16) {
                    ulonglong2 st = make_ulonglong2(0, 0);
                    if (threadIdx.x < nc) {
                        const int prn = de[c0 + threadIdx.x].prn;
                        if (prn > 0 && prn <= 32) {
                            s_prn[threadIdx.x] = prn;
                            st.x = ln_carr_slope_mode(de[c0 + threadIdx.x].carr_step, int_carrier);
                            st.y = ln_code_slope(de[c0 + threadIdx.x].code_step);  // > 0 for every active slot
                        }
                    }
                    s_step[2 * threadIdx.x] = st;
                    s_step[2 * threadIdx.x + 1] = make_ulonglong2(ln_split_dx(st.x), ln_split_dy(st.y));
                }
                const uint32_t in_flight = *s_tx;
                __syncthreads();
                if (in_flight) {  // every thread observes the phase: the copied tables are visible to it afterwards
                    if (!ln_mbar_wait(bar_saddr, bar_phase) && threadIdx.x == 0) atomicExch(err, 0x20000000);
                    bar_phase ^= 1u;
                }
            }

            for (int wb = wb0 + warp; wb < wb1; wb += LN_WARPS) {
                const int tile = wb / WPT;
                const uint32_t mb = (uint32_t) ((wb % WPT) * LN_WB);  // the warp-block's offset inside the tile
                __syncwarp();
                if (lane < nc) {
                    // anchor of the warp-block: tile anchor + mb steps; the chip table's shared address goes into G
                    ulonglong2 a = anch[((size_t) e * ntiles + tile) * C + c0 + lane];
                    const ulonglong2 st = s_step[2 * lane];
                    a.x += (uint64_t) mb * st.x;
                    a.y += (uint64_t) mb * st.y + ((uint64_t) (chip_saddr + (uint32_t) lane * 4 * LN_VS) << LN_GBITS);
                    my_anch[lane] = a;
                }
                __syncwarp();
                uint32_t* dst = out_epoch + (size_t) wb * LN_WB + lane;
                const int nleft = N - (wb * LN_WB + lane);  // this lane's sample j exists iff 32*j < nleft
                int32_t acc[LN_RUN];
#pragma unroll
                for (int j = 0; j < LN_RUN; j++) acc[j] = 0x8000;  // I half biased: it never borrows from the Q half
                if (g > 0) {
#pragma unroll
                    for (int j = 0; j < LN_RUN; j++)
                        if (32 * j < nleft) acc[j] = (int32_t) dst[32 * j];  // raw packed sums of the earlier groups
                }
                // carried shared addresses: step record, warp-block anchor, LUT of the slot
                uint32_t rp = step_saddr, ap = anch_saddr, lut = lut_saddr;
                const uint32_t rp_end = step_saddr + (uint32_t) nc * 32u;
                for (; rp != rp_end; rp += 32u, ap += 16u, lut += 2048u) {
                    const ulonglong2 st = ln_lds_v2(rp);
                    if (st.y == 0) continue;  // inactive slot (uniform)
                    const ulonglong2 sd = ln_lds_v2(rp + 16u);
                    const ulonglong2 a = ln_lds_v2(ap);
                    // the lane's exact start, then the split-word lines (ln_kernel_index is the same arithmetic)
                    uint64_t X = (a.x + (uint64_t) lane * st.x) >> LN_XSH;
                    uint64_t Y = (a.y + (uint64_t) lane * st.y) >> LN_YSH;
#pragma unroll
                    for (int j = 0; j < LN_RUN; j++) {
                        acc[j] += ln_lds_lut(lut, (uint32_t) (X >> 32)) * ln_lds_s8((uint32_t) (Y >> 32));
                        X += sd.x;
                        Y += sd.y;
                    }
                }
                if (g == ngroups - 1) {
#pragma unroll
                    for (int j = 0; j < LN_RUN; j++)
                        if (32 * j < nleft) dst[32 * j] = (uint32_t) acc[j] ^ 0x8000u;  // un-bias: the int16 pair as stored
                } else {
#pragma unroll
                    for (int j = 0; j < LN_RUN; j++)
                        if (32 * j < nleft) dst[32 * j] = (uint32_t) acc[j];
                }
            }
        }
        cur_e = e;
    }
}

}  // namespace gpsiq
