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
// for the rest of the run; those few (tile, channel, run) triples are appended to
// a work list by the prologue, and k_tile_fixup adds (right - wrong) into a dense
// per-sample correction array that the main pass starts its accumulators from.
//
// Per sub-batch of FX_SUB_EPOCHS epochs (sized so that the records stay in L2):
//   k_tile_prologue  1 thread per (epoch, tile, channel, NCO): segment lists
//   k_tile_fixup     1 thread per listed (tile, channel, run): corrections
//   k_synth_fixed    4 tile-workers of 128 threads per CTA: the sample loop  Packed accumulation is
// exact while |sum I|, |sum Q| <= 32767; epochs that could exceed that, or
// whose steps are outside the segment-list contract, are flagged by k_prepare
// and rendered by k_synth_lanes instead.
#pragma once

#define FX_THREADS 128
#define FX_RUN 8
#define FX_TILE (FX_THREADS * FX_RUN)
#define FX_MAXC 32
#define FX_CGROUP 16  // channels whose tables are resident in shared memory at a time

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


// ---- per-tile record written by k_tile_prologue, read by k_synth_fixed ------
// One record per (epoch, tile): the segment lists of the 2C (channel, NCO) tasks,
// the segment index of every 8-sample run, and which (channel, run) pairs contain
// a segment boundary.  Layout (bytes), C = channel slots:
//   segF   [C][FX_SCARR + FX_SCODE] u64   fixed-point state at each segment start
//   segn   [C][FX_SCARR + FX_SCODE] u16   tile-relative sample index of each start
//   runseg [2C][FX_THREADS]         u8    segment index of each run's first sample
//   nseg   [2C]                     u16
// plus, in a separate array zeroed by the host before the prologue,
//   fixmask[C][FX_THREADS/32]       u32   bit r: run r of the channel contains a segment boundary
#define FX_SCARR 48
#define FX_SCODE 24
#define FX_SPER (FX_SCARR + FX_SCODE)

__host__ __device__ inline size_t fx_rec_segF(int C) { (void) C; return 0; }
__host__ __device__ inline size_t fx_rec_segn(int C) { return (size_t) C * FX_SPER * 8; }
__host__ __device__ inline size_t fx_rec_runseg(int C) { return fx_rec_segn(C) + (size_t) C * FX_SPER * 2; }
__host__ __device__ inline size_t fx_rec_nseg(int C) { return fx_rec_runseg(C) + (size_t) 2 * C * FX_THREADS; }
__host__ __device__ inline size_t fx_rec_bytes(int C) { return (fx_rec_nseg(C) + (size_t) 2 * C * 2 + 15) & ~(size_t) 15; }
__host__ __device__ inline size_t fx_fixmask_words(int C) { return (size_t) C * (FX_THREADS / 32); }
// segment storage of task (c, nco) inside a record: carrier lists first in each channel's block
__host__ __device__ inline int fx_seg_base(int c, int is_carrier) { return c * FX_SPER + (is_carrier ? 0 : FX_SCARR); }
__host__ __device__ inline int fx_seg_cap(int is_carrier) { return is_carrier ? FX_SCARR : FX_SCODE; }

// Tile prologue for one (epoch, tile, channel, NCO): walk the tile's segments from
// the exact tile-start state.  One thread per task; a warp holds 32 consecutive
// tiles of the same (epoch, channel, NCO), so its lanes see the same step and
// similar segment statistics.
template <int MODE>
__device__ __forceinline__ void fx_tile_walk(unsigned char* rec, uint32_t* fixmask, uint32_t* work, int* nwork,
                                             int work_cap, uint32_t tile_id, int C, int c, const BinadeTab& tab,
                                             double x, double d, int len, int icode, int kbit, uint64_t navbits,
                                             int* overflow) {
    constexpr int IS_CARR = (MODE == NCO_CARRIER) ? 1 : 0;
    const int task = c * 2 + IS_CARR;
    uint64_t* segF = (uint64_t*) (rec + fx_rec_segF(C)) + fx_seg_base(c, IS_CARR);
    uint16_t* segn = (uint16_t*) (rec + fx_rec_segn(C)) + fx_seg_base(c, IS_CARR);
    uint32_t* runseg = (uint32_t*) (rec + fx_rec_runseg(C) + (size_t) task * FX_THREADS);
    fixmask += c * (FX_THREADS / 32);
    int n = 0, nseg = 0, rnext = 0;
    uint32_t word = 0;
    while (n < len) {
        if (nseg >= fx_seg_cap(IS_CARR)) { *overflow = 1; break; }
        const uint32_t pol = (uint32_t) (navbits >> (kbit & 63)) & 1u;
        segF[nseg] = (MODE == NCO_CARRIER) ? fx_carr_fixed(x) : fx_code_fixed(x, pol);
        segn[nseg] = (uint16_t) n;
        if (n & (FX_RUN - 1)) {  // a boundary inside a run: that (channel, run) needs a fix-up
            const int r = n >> 3;
            const uint32_t bit = 1u << (r & 31);
            if (!(atomicOr(&fixmask[r >> 5], bit) & bit)) {  // first boundary seen in this (channel, run)
                const int slot = atomicAdd(nwork, 1);
                if (slot < work_cap) work[slot] = (tile_id << 12) | ((uint32_t) c << 7) | (uint32_t) r;
                else *overflow = 1;
            }
        }
        const int k = run_in_binade<MODE>(x, tab, len - 1 - n);  // samples n .. n+k share the segment
        {   // runs rnext .. rlast have their first sample in this segment: word-wise fill of the byte map
            const int rlast = (n + k) >> 3;
            if (rlast >= rnext) {
                const uint32_t rep = (uint32_t) nseg * 0x01010101u;
                int r = rnext;
                if (r & 3) {  // finish the partially filled word
                    const int stop = min(rlast + 1, (r | 3) + 1);
                    word |= rep & (((stop & 3) ? ((1u << ((stop & 3) * 8)) - 1u) : 0xffffffffu) & ~((1u << ((r & 3) * 8)) - 1u));
                    r = stop;
                    if (!(r & 3)) { runseg[(r >> 2) - 1] = word; word = 0; }
                }
                for (; r + 4 <= rlast + 1; r += 4) runseg[r >> 2] = rep;  // whole words
                if (r <= rlast) {  // start a new partial word
                    word = rep & ((1u << (((rlast + 1) & 3) * 8)) - 1u);
                    r = rlast + 1;
                }
                rnext = r;
            }
        }
        nseg++;
        n += k + 1;
        if (n < len) {  // true step into the next segment
            int w = 0;
            nco_step<MODE>(x, d, w);
            if (MODE == NCO_CODE && w) {
                if (++icode >= 20) { icode = 0; kbit++; }
            }
        }
    }
    if (rnext & 3) runseg[rnext >> 2] = word;
    ((uint16_t*) (rec + fx_rec_nseg(C)))[task] = (uint16_t) nseg;
}

__global__ void __launch_bounds__(128)
k_tile_prologue(const gpsiq_chan_desc* __restrict__ desc, const BinadeTab* __restrict__ tabs,
                const double* __restrict__ code_ck, const int* __restrict__ wrap_ck,
                const CarrLookup carr,
                const int* __restrict__ amp_sum, int* __restrict__ step_flag, unsigned char* __restrict__ recs,
                uint32_t* __restrict__ fixmasks, uint32_t* __restrict__ work, int* __restrict__ nwork, int work_cap,
                int e0, int E, int C, int N, int ntiles) {
    __shared__ BinadeTab s_tab[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tgroups = (ntiles + 31) / 32;
    int w = blockIdx.x * 4 + warp;  // warp id -> (epoch, task, group of 32 tiles)
    const int tg = w % tgroups; w /= tgroups;
    const int task = w % (2 * C);
    const int e = e0 + w / (2 * C);
    if (e >= e0 + E) return;
    const int c = task >> 1;
    const gpsiq_chan_desc d = desc[(size_t) e * C + c];
    const int t = tg * 32 + lane;
    const size_t rec_bytes = fx_rec_bytes(C);
    const size_t tile_id = (size_t) (e - e0) * ntiles + (t < ntiles ? t : 0);
    unsigned char* rec = recs + tile_id * rec_bytes;
    uint32_t* fixmask = fixmasks + tile_id * fx_fixmask_words(C);
    if (d.prn <= 0 || amp_sum[e] > 32767 || (step_flag[e] & 1)) {
        if (t < ntiles) ((uint16_t*) (rec + fx_rec_nseg(C)))[task] = 0;
        return;
    }
    for (int i = lane; i < (int) (sizeof(BinadeTab) / 4); i += 32)
        ((uint32_t*) &s_tab[warp])[i] = ((const uint32_t*) (tabs + ((size_t) e * C + c) * 2 + (task & 1)))[i];
    __syncwarp();
    if (t >= ntiles) return;
    const int n0 = t * FX_TILE;
    const int len = min(FX_TILE, N - n0);
    const size_t o = ((size_t) e * ntiles + t) * C + c;
    int overflow = 0;
    if (task & 1) {
        const double ph = carr_lookup(carr, e, c, t, FX_TILE, N, C, ntiles);
        fx_tile_walk<NCO_CARRIER>(rec, fixmask, work, nwork, work_cap, (uint32_t) tile_id, C, c, s_tab[warp], ph,
                                  d.carr_step, len, 0, 0, 0, &overflow);
    } else {
        const int wr = wrap_ck[o] + d.ms0 % 20;
        fx_tile_walk<NCO_CODE>(rec, fixmask, work, nwork, work_cap, (uint32_t) tile_id, C, c, s_tab[warp], code_ck[o],
                               d.code_step, len, wr % 20, wr / 20, d.navbits, &overflow);
    }
    if (overflow) atomicOr(&step_flag[e], 2);  // segment list too long: the epoch goes to k_synth_lanes
}

// In-segment fixed-point step of the segment starting at fixed-point state f0.
__device__ __forceinline__ uint64_t fx_seg_step(const BinadeTab& tab, uint64_t f0, int is_carrier) {
    const int bi = is_carrier ? fx_carr_bi(f0) : fx_code_bi(f0);
    if (!((tab.valid >> bi) & 1u)) return 0;
    return (uint64_t) fx_scale_delta(tab.delta[bi], (is_carrier ? 11 : 10) - bi);
}

// k_tile_fixup: one thread per listed (tile, channel, run).  For each sample of the run it
// evaluates what the main pass will add for that channel (the run's start state extrapolated)
// and what is right (following the segment lists), and adds the difference to delta[tile][n].
__global__ void __launch_bounds__(128)
k_tile_fixup(const gpsiq_chan_desc* __restrict__ desc, const int32_t* __restrict__ lutp,
             const BinadeTab* __restrict__ tabs, const unsigned char* __restrict__ recs,
             const int8_t* __restrict__ chips, const uint32_t* __restrict__ work, const int* __restrict__ nwork,
             int work_cap, int32_t* __restrict__ delta, int e0, int C, int N, int ntiles) {
    const int count = min(*nwork, work_cap);
    const size_t rec_bytes = fx_rec_bytes(C);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const uint32_t item = work[i];
        const int r = item & 127, c = (item >> 7) & 31;
        const uint32_t tile_id = item >> 12;
        const int e = e0 + tile_id / ntiles, t = tile_id % ntiles;
        const int len = min(FX_TILE, N - t * FX_TILE);
        const unsigned char* rec = recs + (size_t) tile_id * rec_bytes;
        const uint64_t* segF = (const uint64_t*) (rec + fx_rec_segF(C));
        const uint16_t* segn = (const uint16_t*) (rec + fx_rec_segn(C));
        const uint16_t* nsegs = (const uint16_t*) (rec + fx_rec_nseg(C));
        const BinadeTab* tab = tabs + ((size_t) e * C + c) * 2;  // [0] code, [1] carrier
        const int32_t* lut = lutp + ((size_t) e * C + c) * 512;
        const int8_t* chip = chips + (size_t) desc[(size_t) e * C + c].prn * 2048;
        // run-start states, extrapolated exactly as the main pass does
        int ic = fx_seg_base(c, 1) + rec[fx_rec_runseg(C) + (c * 2 + 1) * FX_THREADS + r];
        int ig = fx_seg_base(c, 0) + rec[fx_rec_runseg(C) + (c * 2) * FX_THREADS + r];
        const int endc = fx_seg_base(c, 1) + nsegs[c * 2 + 1], endg = fx_seg_base(c, 0) + nsegs[c * 2];
        uint64_t df = fx_seg_step(tab[1], segF[ic], 1), dg = fx_seg_step(tab[0], segF[ig], 0);
        uint64_t f = segF[ic] + (uint64_t) (r * FX_RUN - segn[ic]) * df;
        uint64_t g = segF[ig] + (uint64_t) (r * FX_RUN - segn[ig]) * dg;
        uint64_t fe = f, dfe = df, ge = g, dge = dg;  // exact: follows the segment lists
        for (int j = 0; j < FX_RUN; j++) {
            const int n = r * FX_RUN + j;
            if (n >= len) break;
            if (ic + 1 < endc && segn[ic + 1] == n) { ic++; fe = segF[ic]; dfe = fx_seg_step(tab[1], fe, 1); }
            if (ig + 1 < endg && segn[ig + 1] == n) { ig++; ge = segF[ig]; dge = fx_seg_step(tab[0], ge, 0); }
            const int ir = (int) (fe >> 55), iw = (int) (f >> 55), cr = (int) (ge >> 53), cw = (int) (g >> 53);
            if (ir != iw || cr != cw) {
                const int32_t right = lut[ir] * (int32_t) chip[cr];
                const int32_t wrong = lut[iw] * (int32_t) chip[cw];
                if (right != wrong) atomicAdd(&delta[(size_t) tile_id * FX_TILE + n], right - wrong);
            }
            f += df; g += dg; fe += dfe; ge += dge;
        }
    }
}

// ---- k_synth_fixed -------------------------------------------------------------
#define FX_WORKERS 4  // tile-workers (128 threads each) per CTA, sharing the epoch's tables
#define FX_TILES_PER_WORKER 2
#define FX_TILES_PER_CTA (FX_WORKERS * FX_TILES_PER_WORKER)

__host__ __device__ inline size_t fx_smem_bytes(int C) {
    const int CG = C < FX_CGROUP ? C : FX_CGROUP;
    return (size_t) CG * 512 * 4 + (size_t) CG * 2048 + (size_t) CG * NBINADE * 16 + FX_WORKERS * fx_rec_bytes(C) + 64;
}

__device__ __forceinline__ void fx_worker_sync(int worker) {
    asm volatile("bar.sync %0, %1;" ::"r"(worker + 1), "n"(FX_THREADS) : "memory");
}

// Fixed-point state of one NCO of channel c at the first sample of run r, and its
// in-segment step: the run's segment start extrapolated inside its binade.
__device__ __forceinline__ void fx_run_state(const unsigned char* rec, const int64_t* dtab, int C, int c, int is_carrier,
                                             int r, uint64_t& f, uint64_t& df) {
    const int task = c * 2 + is_carrier;
    const int slot = fx_seg_base(c, is_carrier) + rec[fx_rec_runseg(C) + task * FX_THREADS + r];
    const uint64_t f0 = ((const uint64_t*) (rec + fx_rec_segF(C)))[slot];
    const int ns = ((const uint16_t*) (rec + fx_rec_segn(C)))[slot];
    df = (uint64_t) dtab[c * NBINADE + (is_carrier ? fx_carr_bi(f0) : fx_code_bi(f0))];
    f = f0 + (uint64_t) (r * FX_RUN - ns) * df;
}

__global__ void __launch_bounds__(FX_WORKERS * FX_THREADS)
k_synth_fixed(const gpsiq_chan_desc* __restrict__ desc, const int32_t* __restrict__ lutp,
              const BinadeTab* __restrict__ tabs, const unsigned char* __restrict__ recs,
              const int32_t* __restrict__ delta, const int8_t* __restrict__ chips, const int* __restrict__ amp_sum,
              const int* __restrict__ step_flag, int16_t* __restrict__ iq, int e0, int C, int N, int ntiles,
              int groups) {
    extern __shared__ __align__(16) unsigned char fx_raw[];
    const size_t rec_bytes = fx_rec_bytes(C);
    const int CG = C < FX_CGROUP ? C : FX_CGROUP;          // channels per table group
    const int ngroups = (C + CG - 1) / CG;
    unsigned char* s_rec0 = fx_raw;
    int64_t* s_dF = (int64_t*) (fx_raw + FX_WORKERS * rec_bytes);
    int64_t* s_dG = s_dF + CG * NBINADE;
    int32_t* s_lut = (int32_t*) (s_dG + CG * NBINADE);
    int8_t* s_chip = (int8_t*) (s_lut + CG * 512);

    const int e = e0 + blockIdx.x / groups;
    const int grp = blockIdx.x % groups;
    if (amp_sum[e] > 32767 || step_flag[e]) return;  // this epoch is rendered by k_synth_lanes
    const gpsiq_chan_desc* de = desc + (size_t) e * C;
    const int tid_cta = threadIdx.x;
    const int worker = tid_cta >> 7, tid = tid_cta & (FX_THREADS - 1);

    // stage the tables of channel group g (amplitude LUTs, chip signs, fixed-point steps)
    auto stage_tables = [&](int g) {
        const int c0 = g * CG, nc = min(CG, C - c0);
        for (int i = tid_cta; i < nc * 512; i += FX_WORKERS * FX_THREADS) {
            const int prn = de[c0 + (i >> 9)].prn;
            s_lut[i] = (prn > 0) ? lutp[((size_t) e * C + c0) * 512 + i] : 0;
            ((uint32_t*) s_chip)[i] = (prn > 0 && prn <= 32) ? ((const uint32_t*) chips)[prn * 512 + (i & 511)] : 0x01010101u;
        }
        for (int i = tid_cta; i < nc * NBINADE; i += FX_WORKERS * FX_THREADS) {
            const int c = i / NBINADE, bi = i - c * NBINADE;
            const BinadeTab* tc = tabs + ((size_t) e * C + c0 + c) * 2;
            s_dG[i] = ((tc[0].valid >> bi) & 1u) ? fx_scale_delta(tc[0].delta[bi], 10 - bi) : 0;
            s_dF[i] = ((tc[1].valid >> bi) & 1u) ? fx_scale_delta(tc[1].delta[bi], 11 - bi) : 0;
        }
    };
    if (ngroups == 1) {  // the usual case: everything resident for the CTA's whole life
        stage_tables(0);
        __syncthreads();
    }

    unsigned char* s_rec = s_rec0 + worker * rec_bytes;
    uint32_t* out_epoch = reinterpret_cast<uint32_t*>(iq) + (size_t) e * N;
    for (int tt = 0; tt < FX_TILES_PER_WORKER; tt++) {
        const int t = grp * FX_TILES_PER_CTA + tt * FX_WORKERS + worker;
        const bool active = t < ntiles;                  // (with several groups every worker keeps hitting the CTA barriers)
        if (!active && ngroups == 1) break;
        const int n0 = t * FX_TILE;
        const int len = active ? min(FX_TILE, N - n0) : 0;
        const size_t tile_id = (size_t) (e - e0) * ntiles + (active ? t : 0);
        const bool mine = active && tid * FX_RUN < len;

        // ---- this worker's tile record -> shared memory
        if (active) {
            const uint4* src = (const uint4*) (recs + tile_id * rec_bytes);
            uint4* dst = (uint4*) s_rec;
            for (int i = tid; i < (int) (rec_bytes / 16); i += FX_THREADS) dst[i] = src[i];
        }
        fx_worker_sync(worker);

        int32_t acc[FX_RUN];
#pragma unroll
        for (int j = 0; j < FX_RUN; j++) acc[j] = 0;
        if (mine) {  // corrections for runs that contain a segment boundary (zero elsewhere)
            const int4* dp = reinterpret_cast<const int4*>(delta + tile_id * FX_TILE + tid * FX_RUN);
            const int4 a = dp[0], b = dp[1];
            acc[0] = a.x; acc[1] = a.y; acc[2] = a.z; acc[3] = a.w;
            acc[4] = b.x; acc[5] = b.y; acc[6] = b.z; acc[7] = b.w;
        }
        const uint16_t* nseg = (const uint16_t*) (s_rec + fx_rec_nseg(C));
        for (int g = 0; g < ngroups; g++) {
            if (ngroups > 1) {
                __syncthreads();   // everyone is done with the previous group's tables
                stage_tables(g);
                __syncthreads();
            }
            if (!mine) continue;
            const int c0 = g * CG, nc = min(CG, C - c0);
            // ---- main pass: uniform integer inner loop
            for (int cl = 0; cl < nc; cl++) {
                const int c = c0 + cl;
                if (nseg[c * 2] == 0) continue;  // inactive slot (uniform across the worker)
                uint64_t f, df, g2, dg;
                fx_run_state(s_rec, s_dF - (size_t) c0 * NBINADE, C, c, 1, tid, f, df);
                fx_run_state(s_rec, s_dG - (size_t) c0 * NBINADE, C, c, 0, tid, g2, dg);
                const int32_t* lut = s_lut + cl * 512;
                const int8_t* chip = s_chip + cl * 2048;
#pragma unroll
                for (int j = 0; j < FX_RUN; j++) {
                    acc[j] += lut[(uint32_t) (f >> 55)] * (int32_t) chip[(uint32_t) (g2 >> 53)];
                    f += df;
                    g2 += dg;
                }
            }
        }
        if (mine) {
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
        fx_worker_sync(worker);  // the record buffer is reused for the next tile
    }
}

}  // namespace gpsiq
