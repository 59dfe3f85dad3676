// synth_fixed.cuh — k_synth_fixed: the production synthesis kernel.
//
// Replaces the reference's per-sample loop (plutogpssim.c:2689-2756) with a
// uniform, branch-free integer inner loop:
//
//   thread  = a run of FX_RUN = 8 consecutive samples (32 B of output, stored as
//             two 128-bit stores; a warp writes 1 KB contiguous)
//   CTA     = 128 threads = one tile of 1024 samples at a time; a CTA walks
//             FX_TILES_PER_CTA consecutive tiles of ONE epoch so that the tables
//             staged in shared memory are reused
//   smem    = per channel: amplitude LUT (512 x packed int16 I,Q), C/A chips as
//             +-1 bytes (both NAV polarities), per-binade fixed-point NCO steps,
//             per-run NCO start states, per-tile segment lists
//
// Why integers are exact here: inside one binade the reference's binary64 NCO
// is a fixed-point NCO (nco_scan.cuh).  A tile prologue (one lane per
// (channel, NCO)) walks the tile's segments from the exact tile-start state
// and leaves, for every 8-sample run, the 64-bit fixed-point state at its
// first sample:
//   carrier  F = phase * 2^64        -> table index  = F >> 55      (plutogpssim.c:2697)
//   code     G = chips * 2^53 | nav<<63 -> chip+polarity = G >> 53  (plutogpssim.c:2737, 2732)
// The main pass then does, per channel and sample: two 64-bit adds, two shifts,
// two shared-memory loads and one multiply-add on the packed I/Q word
// (plutogpssim.c:2701-2706), with no branch.  A run that contains a segment
// boundary (binade crossing, wrap, NAV bit edge) extrapolates the wrong segment
// for the rest of the run; those few (channel, run) pairs are listed by the
// prologue and a fix-up pass adds (right - wrong) to a per-sample correction
// that the main pass starts its accumulators from.  Packed accumulation is
// exact while |sum I|, |sum Q| <= 32767; epochs that could exceed that, or
// whose steps are outside the segment-list contract, are flagged by k_prepare
// and rendered by k_synth_lanes instead.
#pragma once

#define FX_THREADS 128
#define FX_RUN 8
#define FX_TILE (FX_THREADS * FX_RUN)
#define FX_SEGMAX 96
#define FX_MAXC 24
#define FX_TILES_PER_CTA 8

namespace gpsiq {

constexpr uint64_t FX_POL = 1ULL << 63;

// ---- binary64 state -> 64-bit fixed point --------------------------------
__device__ __forceinline__ uint64_t fx_carr_fixed(double x) {
    const int64_t b = f64_bits(x);
    const int e = (int) (b >> 52);
    if (e <= 0) return 0;
    const uint64_t m = ((uint64_t) b & 0xfffffffffffffULL) | (1ULL << 52);
    const int sh = e - 1011;
    if (sh >= 12) return ~0ULL;  // phase == 1.0 (plutogpssim.c:2745-2746 can round to it): index 511, as k_synth_lanes
    if (sh >= 0) return m << sh;
    return (sh > -64) ? (m >> (-sh)) : 0;
}
__device__ __forceinline__ uint64_t fx_code_fixed(double x, uint32_t pol) {
    const int64_t b = f64_bits(x);
    const int e = (int) (b >> 52);
    uint64_t g = 0;
    if (e > 0) {
        const uint64_t m = ((uint64_t) b & 0xfffffffffffffULL) | (1ULL << 52);
        const int sh = e - 1022;
        g = (sh >= 0) ? (m << sh) : ((sh > -64) ? (m >> (-sh)) : 0);
    }
    return g | ((uint64_t) pol << 63);
}
__device__ __forceinline__ int64_t fx_scale_delta(int64_t delta, int sh) {
    return (sh >= 0) ? (int64_t) ((uint64_t) delta << sh) : (delta >> (-sh));
}
__device__ __forceinline__ int fx_carr_bi(uint64_t f) { return min(__clzll((long long) f), NBINADE - 1); }
__device__ __forceinline__ int fx_code_bi(uint64_t g) { return min(__clzll((long long) (g & ~FX_POL)) - 1, NBINADE - 1); }

struct FxSmem {
    int C;
    int32_t* lut;      // [C][512]  packed (Q << 16) + I
    int8_t* chip;      // [C][2048] +-1, index = nav polarity << 10 | chip
    uint64_t* slotF;   // [C][FX_THREADS]
    uint64_t* slotG;   // [C][FX_THREADS]
    int64_t* dF;       // [C][NBINADE]
    int64_t* dG;       // [C][NBINADE]
    BinadeTab* tab;    // [C][2]
    uint64_t* segF;    // [2C][FX_SEGMAX]
    int16_t* segn;     // [2C][FX_SEGMAX]
    int* nseg;         // [2C]
    int32_t* delta;    // [FX_TILE]
    uint32_t* fixmask; // [C][FX_THREADS/32]
    uint16_t* work;    // [C * FX_THREADS]
    int* nwork;
};

__host__ __device__ inline size_t fx_smem_bytes(int C) {
    size_t b = 0;
    b += (size_t) C * 512 * 4;                 // lut
    b += (size_t) C * 2048;                    // chip
    b += (size_t) C * FX_THREADS * 8 * 2;      // slotF, slotG
    b += (size_t) C * NBINADE * 8 * 2;         // dF, dG
    b += (size_t) C * 2 * sizeof(BinadeTab);   // tab
    b += (size_t) C * 2 * FX_SEGMAX * 8;       // segF
    b += (size_t) C * 2 * FX_SEGMAX * 2;       // segn
    b += (size_t) C * 2 * 4;                   // nseg
    b += (size_t) FX_TILE * 4;                 // delta
    b += (size_t) C * (FX_THREADS / 32) * 4;   // fixmask
    b += (size_t) C * FX_THREADS * 2;          // work
    b += 16;                                   // nwork + pad
    return b + 64;
}

__device__ __forceinline__ void fx_carve(FxSmem& s, unsigned char* base, int C) {
    // 8-byte members first
    s.C = C;
    s.slotF = (uint64_t*) base;               base += (size_t) C * FX_THREADS * 8;
    s.slotG = (uint64_t*) base;               base += (size_t) C * FX_THREADS * 8;
    s.dF = (int64_t*) base;                   base += (size_t) C * NBINADE * 8;
    s.dG = (int64_t*) base;                   base += (size_t) C * NBINADE * 8;
    s.tab = (BinadeTab*) base;                base += (size_t) C * 2 * sizeof(BinadeTab);
    s.segF = (uint64_t*) base;                base += (size_t) C * 2 * FX_SEGMAX * 8;
    s.lut = (int32_t*) base;                  base += (size_t) C * 512 * 4;
    s.delta = (int32_t*) base;                base += (size_t) FX_TILE * 4;
    s.nseg = (int*) base;                     base += (size_t) C * 2 * 4;
    s.fixmask = (uint32_t*) base;             base += (size_t) C * (FX_THREADS / 32) * 4;
    s.nwork = (int*) base;                    base += 16;
    s.segn = (int16_t*) base;                 base += (size_t) C * 2 * FX_SEGMAX * 2;
    s.work = (uint16_t*) base;                base += (size_t) C * FX_THREADS * 2;
    s.chip = (int8_t*) base;
}

// Contribution of channel c to sample j of run r, extrapolating the run's start state:
// exactly what the main pass adds (shared by the fix-up pass for the "wrong" value).
__device__ __forceinline__ int32_t fx_extrapolated(const FxSmem& s, int c, int r, int j) {
    const uint64_t f0 = s.slotF[c * FX_THREADS + r], g0 = s.slotG[c * FX_THREADS + r];
    const uint64_t f = f0 + (uint64_t) j * (uint64_t) s.dF[c * NBINADE + fx_carr_bi(f0)];
    const uint64_t g = g0 + (uint64_t) j * (uint64_t) s.dG[c * NBINADE + fx_code_bi(g0)];
    return s.lut[c * 512 + (int) (f >> 55)] * (int32_t) s.chip[c * 2048 + (int) (g >> 53)];
}

// Exact fixed-point state of one NCO at tile sample n, from its segment list.
__device__ __forceinline__ uint64_t fx_exact(const FxSmem& s, int task, int n, const int64_t* dtab, bool code) {
    const int16_t* sn = s.segn + task * FX_SEGMAX;
    int lo = 0, hi = s.nseg[task] - 1;
    while (lo < hi) {  // last segment with start <= n
        const int mid = (lo + hi + 1) >> 1;
        if (sn[mid] <= n) lo = mid; else hi = mid - 1;
    }
    const uint64_t f0 = s.segF[task * FX_SEGMAX + lo];
    const int bi = code ? fx_code_bi(f0) : fx_carr_bi(f0);
    return f0 + (uint64_t) (n - sn[lo]) * (uint64_t) dtab[bi];
}

// Tile prologue for one (channel, NCO): segment list + per-run start states.
template <int MODE>
__device__ void fx_prologue(FxSmem& s, int c, double x, double d, int len, int icode, int kbit, uint64_t navbits,
                            int* err) {
    const int task = c * 2 + (MODE == NCO_CARRIER ? 1 : 0);
    const BinadeTab& tab = s.tab[task];
    const int64_t* dtab = (MODE == NCO_CARRIER ? s.dF : s.dG) + c * NBINADE;
    uint64_t* slot = (MODE == NCO_CARRIER ? s.slotF : s.slotG) + c * FX_THREADS;
    uint64_t* segF = s.segF + task * FX_SEGMAX;
    int16_t* segn = s.segn + task * FX_SEGMAX;
    int n = 0, nseg = 0;
    while (n < len) {
        if (nseg >= FX_SEGMAX) { atomicExch(err, 0x40000000 | task); break; }  // outside the step contract
        const uint32_t pol = (uint32_t) (navbits >> (kbit & 63)) & 1u;
        const uint64_t f0 = (MODE == NCO_CARRIER) ? fx_carr_fixed(x) : fx_code_fixed(x, pol);
        segn[nseg] = (int16_t) n;
        segF[nseg] = f0;
        nseg++;
        if (n & (FX_RUN - 1)) {  // a boundary inside a run: that (channel, run) needs a fix-up
            const int r = n >> 3;
            const uint32_t bit = 1u << (r & 31);
            const uint32_t old = atomicOr(&s.fixmask[c * (FX_THREADS / 32) + (r >> 5)], bit);
            if (!(old & bit)) s.work[atomicAdd(s.nwork, 1)] = (uint16_t) (c * FX_THREADS + r);
        }
        const int k = run_in_binade<MODE>(x, tab, len - 1 - n);  // samples n .. n+k share the segment
        const int64_t df = dtab[(MODE == NCO_CARRIER) ? fx_carr_bi(f0) : fx_code_bi(f0)];
        for (int r = (n + FX_RUN - 1) >> 3; (r << 3) <= n + k; r++)
            slot[r] = f0 + (uint64_t) ((r << 3) - n) * (uint64_t) df;
        n += k + 1;
        if (n < len) {  // true step into the next segment
            int w = 0;
            nco_step<MODE>(x, d, w);
            if (MODE == NCO_CODE && w) {
                if (++icode >= 20) { icode = 0; kbit++; }
            }
        }
    }
    s.nseg[task] = nseg;
}

__global__ void __launch_bounds__(FX_THREADS)
k_synth_fixed(const gpsiq_chan_desc* __restrict__ desc, const int32_t* __restrict__ lutp,
              const BinadeTab* __restrict__ tabs, const double* __restrict__ code_ck,
              const int* __restrict__ wrap_ck, const double* __restrict__ carr_ck, size_t ck_plane,
              const CarrInfo* __restrict__ info, const int8_t* __restrict__ chips, const int* __restrict__ amp_sum,
              const int* __restrict__ step_flag, int16_t* __restrict__ iq, int* __restrict__ err, int C, int N,
              int ntiles, int groups) {
    extern __shared__ __align__(16) unsigned char fx_raw[];
    FxSmem s;
    fx_carve(s, fx_raw, C);
    const int e = blockIdx.x / groups;
    const int grp = blockIdx.x - e * groups;
    if (amp_sum[e] > 32767 || step_flag[e]) return;  // this epoch is rendered by k_synth_lanes
    const gpsiq_chan_desc* de = desc + (size_t) e * C;
    const int tid = threadIdx.x;

    // ---- stage the epoch's tables (reused for FX_TILES_PER_CTA tiles)
    for (int i = tid; i < C * 512; i += FX_THREADS) s.lut[i] = (de[i >> 9].prn > 0) ? lutp[(size_t) e * C * 512 + i] : 0;
    for (int i = tid; i < C * 512; i += FX_THREADS) {  // 2048 bytes per channel, 4 at a time
        const int c = i >> 9;
        const int prn = de[c].prn;
        ((uint32_t*) s.chip)[i] = (prn > 0 && prn <= 32) ? ((const uint32_t*) chips)[prn * 512 + (i & 511)] : 0x01010101u;
    }
    for (int i = tid; i < C * 2 * (int) (sizeof(BinadeTab) / 4); i += FX_THREADS)
        ((uint32_t*) s.tab)[i] = ((const uint32_t*) (tabs + (size_t) e * C * 2))[i];
    __syncthreads();
    for (int i = tid; i < C * NBINADE; i += FX_THREADS) {
        const int c = i / NBINADE, bi = i - c * NBINADE;
        const BinadeTab& tc = s.tab[c * 2], &tp = s.tab[c * 2 + 1];
        s.dG[i] = ((tc.valid >> bi) & 1u) ? fx_scale_delta(tc.delta[bi], 10 - bi) : 0;
        s.dF[i] = ((tp.valid >> bi) & 1u) ? fx_scale_delta(tp.delta[bi], 11 - bi) : 0;
    }

    uint32_t* out_epoch = reinterpret_cast<uint32_t*>(iq) + (size_t) e * N;
    for (int tt = 0; tt < FX_TILES_PER_CTA; tt++) {
        const int t = grp * FX_TILES_PER_CTA + tt;
        if (t >= ntiles) break;
        const int n0 = t * FX_TILE;
        const int len = min(FX_TILE, N - n0);

        // ---- reset per-tile scratch
        for (int i = tid; i < FX_TILE; i += FX_THREADS) s.delta[i] = 0;
        for (int i = tid; i < C * (FX_THREADS / 32); i += FX_THREADS) s.fixmask[i] = 0;
        if (tid == 0) *s.nwork = 0;
        __syncthreads();

        // ---- prologue: task = (channel, NCO); tasks are spread over the 4 warps
        {
            const int warp = tid >> 5, lane = tid & 31;
            const int task = lane * 4 + warp;
            if (task < 2 * C) {
                const int c = task >> 1;
                const gpsiq_chan_desc d = de[c];
                if (d.prn > 0) {
                    const size_t o = ((size_t) e * ntiles + t) * C + c;
                    if (task & 1) {
                        const CarrInfo inf = info[(size_t) e * C + c];
                        double ph;
                        if (n0 < inf.n1 || inf.n1 >= N) ph = carr_ck[o];
                        else ph = __dadd_rn(carr_ck[(size_t) inf.variant * ck_plane + o], inf.delta);
                        fx_prologue<NCO_CARRIER>(s, c, ph, d.carr_step, len, 0, 0, 0, err);
                    } else {
                        const int w = wrap_ck[o] + d.ms0 % 20;
                        fx_prologue<NCO_CODE>(s, c, code_ck[o], d.code_step, len, w % 20, w / 20, d.navbits, err);
                    }
                } else {
                    s.nseg[task] = 0;
                }
            }
        }
        __syncthreads();

        // ---- fix-up pass: runs that contain a segment boundary
        for (int i = tid; i < *s.nwork; i += FX_THREADS) {
            const int c = s.work[i] / FX_THREADS, r = s.work[i] % FX_THREADS;
            for (int j = 0; j < FX_RUN; j++) {
                const int n = r * FX_RUN + j;
                if (n >= len) break;
                const uint64_t f = fx_exact(s, c * 2 + 1, n, s.dF + c * NBINADE, false);
                const uint64_t g = fx_exact(s, c * 2, n, s.dG + c * NBINADE, true);
                const int32_t right = s.lut[c * 512 + (int) (f >> 55)] * (int32_t) s.chip[c * 2048 + (int) (g >> 53)];
                const int32_t wrong = fx_extrapolated(s, c, r, j);
                if (right != wrong) atomicAdd(&s.delta[n], right - wrong);
            }
        }
        __syncthreads();

        // ---- main pass: uniform integer inner loop
        if (tid * FX_RUN < len) {
            int32_t acc[FX_RUN];
            {
                const int4 a = *reinterpret_cast<const int4*>(s.delta + tid * FX_RUN);
                const int4 b = *reinterpret_cast<const int4*>(s.delta + tid * FX_RUN + 4);
                acc[0] = a.x; acc[1] = a.y; acc[2] = a.z; acc[3] = a.w;
                acc[4] = b.x; acc[5] = b.y; acc[6] = b.z; acc[7] = b.w;
            }
            for (int c = 0; c < C; c++) {
                if (s.nseg[c * 2] == 0) continue;  // inactive slot (uniform across the CTA)
                uint64_t f = s.slotF[c * FX_THREADS + tid], g = s.slotG[c * FX_THREADS + tid];
                const uint64_t df = (uint64_t) s.dF[c * NBINADE + fx_carr_bi(f)];
                const uint64_t dg = (uint64_t) s.dG[c * NBINADE + fx_code_bi(g)];
                const int32_t* lut = s.lut + c * 512;
                const int8_t* chip = s.chip + c * 2048;
#pragma unroll
                for (int j = 0; j < FX_RUN; j++) {
                    acc[j] += lut[(uint32_t) (f >> 55)] * (int32_t) chip[(uint32_t) (g >> 53)];
                    f += df;
                    g += dg;
                }
            }
            // packed (Q<<16)+I with signed I  ->  int16 pair (plutogpssim.c:2754-2755)
            uint32_t w[FX_RUN];
#pragma unroll
            for (int j = 0; j < FX_RUN; j++) w[j] = (uint32_t) acc[j] + (((uint32_t) acc[j] & 0x8000u) << 1);
            uint32_t* dst = out_epoch + n0 + tid * FX_RUN;
            if (tid * FX_RUN + FX_RUN <= len) {
                *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<uint4*>(dst + 4) = make_uint4(w[4], w[5], w[6], w[7]);
            } else {
                for (int j = 0; j < FX_RUN && tid * FX_RUN + j < len; j++) dst[j] = w[j];
            }
        }
        __syncthreads();
    }
}

}  // namespace gpsiq
