// gpsiq.cu — sm_100a kernels + C-ABI (include/gpsiq.h) of the GPS L1 C/A
// baseband I/Q synthesizer.  Replaces the reference's per-sample loop,
// /root/reference/plutogpssim.c:2689-2756, for whole batches of 0.1 s epochs.
//
// Pipeline for one batch of E epochs x C channel slots x N samples/epoch
// (all on one stream, no host synchronisation in between):
//
//   k_prepare       (E*C CTAs)   per-(epoch,slot) amplitude LUT
//                                lut[k] = ((int)(cos[k]*gain), (int)(sin[k]*gain))
//                                == plutogpssim.c:2701-2702 with the +-1 BPSK sign
//                                factored out (trunc is odd-symmetric)
//   k_scan_code     (E*C threads) exact code-NCO state at every tile boundary
//                                (restarts every epoch: plutogpssim.c:1770)
//   k_scan_carrier  (C threads)  exact carrier-NCO state at every tile boundary;
//                                serial over the epochs of the batch because
//                                carr_phase chains across epochs (plutogpssim.c:2741)
//   k_synth_*       (E*tiles)    the per-sample work: chip lookup, NAV bit,
//                                carrier LUT, across-channel accumulate, int16 pack
//
// See DESIGN.md for the data layout and the roofline of each kernel.
#include <cuda.h>          // driver API types only: entry points come from cudaGetDriverEntryPoint (no libcuda link)
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../include/gpsiq.h"
#include "../../include/gpsiq_desc.h"
#include "nco_scan.cuh"
#include "synth_line.cuh"

using namespace gpsiq;

// ---------------------------------------------------------------------------
// tables
// ---------------------------------------------------------------------------
#include "carrier_tables.inc"  // static const int16_t k_sin512[512], k_cos512[512] (generated, tools/gen_tables.py)

__constant__ int16_t c_sin512[512];
__constant__ int16_t c_cos512[512];

// C/A code bitmaps: 32 words per PRN, bit (i&31) of word (i>>5) = chip i.
#define CA_WORDS 32

// IS-GPS-200 Table 3-I code phase selection: G2i = stage a XOR stage b of G2.
static const unsigned char k_g2_taps[32][2] = {
    {2, 6}, {3, 7}, {4, 8}, {5, 9}, {1, 9}, {2, 10}, {1, 8}, {2, 9}, {3, 10}, {2, 3}, {3, 4},
    {5, 6}, {6, 7}, {7, 8}, {8, 9}, {9, 10}, {1, 4}, {2, 5}, {3, 6}, {4, 7}, {5, 8}, {6, 9},
    {1, 3}, {4, 6}, {5, 7}, {6, 8}, {7, 9}, {8, 10}, {1, 6}, {2, 7}, {3, 8}, {4, 9}};

// Same sequence as the reference's codegen (plutogpssim.c:207-244), which uses
// the equivalent "G2 delay" formulation; tests compare the two chip for chip.
static void ca_generate(int prn, uint8_t* chips) {
    unsigned g1 = 0x3ff, g2 = 0x3ff;  // bit s-1 = stage s
    const int a = k_g2_taps[prn - 1][0] - 1, b = k_g2_taps[prn - 1][1] - 1;
    for (int i = 0; i < GPSIQ_CA_LEN; i++) {
        unsigned o1 = (g1 >> 9) & 1u;
        unsigned o2 = ((g2 >> a) ^ (g2 >> b)) & 1u;
        chips[i] = (uint8_t) (o1 ^ o2);
        unsigned f1 = ((g1 >> 2) ^ (g1 >> 9)) & 1u;
        unsigned f2 = ((g2 >> 1) ^ (g2 >> 2) ^ (g2 >> 5) ^ (g2 >> 7) ^ (g2 >> 8) ^ (g2 >> 9)) & 1u;
        g1 = ((g1 << 1) | f1) & 0x3ff;
        g2 = ((g2 << 1) | f2) & 0x3ff;
    }
}

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
#define TIMING_RING 64

// Optional timeline trace (GPSIQ_TRACE=1 in the environment): an event after every kernel launch,
// dumped by gpsiq_trace_dump as milliseconds since the first one.  Diagnostics only.
#define TRACE_MAX 4096
struct TraceRec { cudaEvent_t ev; const char* label; int stream_id; double host_ms; };   // host_ms: when the host enqueued it


// Everything the scan phases produce for one batch and the render phase consumes.  There are two
// sets so that gpsiq_submit_device can scan batch k+1 while gpsiq_fetch_device renders batch k;
// the working pointers in gpsiq_ctx (d_lut, d_carr_ck, ...) are switched to one set before enqueuing.
// (see k_carr_final)
struct Handoff {
    const unsigned long long* in_flag;   // own mailbox flag (NULL: start from carr_state)
    unsigned long long in_seq;
    const double* in_slot;               // own mailbox slot holding message in_seq
    double* out_slot;                    // next GPU's mailbox slot for message out_seq (NULL: no send)
    unsigned long long* out_flag;        // next GPU's flag
    unsigned long long out_seq;
    unsigned int* counter;               // own: slots that have stored their end phase (reset by the last one)
    double* start_copy;                  // optional: the exact start phases, for the caller's estimate feedback
    int* err;
};
struct SliceRes;
struct ScanSet {
    gpsiq_chan_desc* d_descbuf;
    int2* d_lut; int32_t* d_lutp; int* d_flags; double* d_code_ck; int* d_wrap_ck; double* d_carr_ck;
    double* d_drift; CarrSpec* d_spec; CarrSpec* d_specE; ChunkInfo* d_cinfo; CarrInfo* d_info;
    CarrSpec* d_specG; GroupInfo* d_ginfo; double* d_traceG;
    CarrSpec* d_specS; double* d_startS; SliceRes* d_sres;   // level 5: slice-level speculation, group entry phases, match result
    TieEvent* d_tieG; TieEvent* d_tieS;                      // first tie-wrap of every group- / slice-level trajectory
    double* d_start0;             // [C] exact phases at the first sample of the batch (once its chain has run)
    double* d_rate_used;          // [C] the residual-rate estimate the batch was prepared with (k_bias_update)
    double* d_ccum_used;          // [C] free-running estimates: the total correction the batch's start estimate had received
    double* d_adv; double* d_carr_trace; uint32_t* d_ustart;
    LineEpoch* d_lrecs; uint32_t* d_elist; uint32_t* d_hazlist; int* d_line_counters; LinePatch* d_patches;
    int anchored;                 // the batch's tile anchors, safety check and patch list have been enqueued
    cudaEvent_t anch_ready, anchor_done;   // anchors written (the sample kernel may start) / patch list complete
    double* d_est;                // [C] estimated phases at the first sample of the batch (pipelined submits)
    double* d_exact_end;          // [C] exact phases after the batch's last sample (once its chain has run)
    cudaStream_t stream;          // pipelined submits: each set scans on its own stream, so that the scans of
                                  // consecutive batches overlap (only the exact chain is ordered batch after batch)
    cudaEvent_t scan_done, render_done, spec_done, adv_done, est_done;
    cudaEvent_t spec_all;         // every speculative level of the batch has run (the chain may be enqueued on another stream)
    long long seq;                // number of the batch the set holds (-1: none yet)
    const gpsiq_chan_desc* desc;  // the batch's descriptors (device)
    int n_epochs;
    int phase;                    // 0 free, 1 prepared, 2 speculated, 3 chained (waiting to be rendered)
};

#define NSETS 4   // scan sets: gpsiq_submit* may run NSETS - 1 batches ahead of the one being rendered

struct gpsiq_ctx {
    gpsiq_config cfg;
    ScanSet sets[NSETS];
    // the ring of scan sets, one cursor per phase: a batch is prepared (set_prep), speculated (set_spec), chained (set_wr)
    // and rendered (set_rd), in that order; the phase API of time-sliced runs may prepare / speculate batches ahead of
    // the one being chained (on another stream), the pipelined submits move the first three cursors together
    int set_prep, set_spec, set_wr, set_rd, set_pending;
    int set_cur;                      // set the working pointers currently point at
    cudaStream_t scan_stream;         // gpsiq_submit_device scans here, ahead of the caller's render stream
    cudaStream_t aux2_stream;         // its code-NCO scan (aux_stream is busy with the tile prologues of the batch being rendered)
    int C, N, T, ntiles, E;
    int sm_count;
    cudaStream_t stream;
    cudaStream_t copy_stream;        // device-to-host copies of finished sub-batches overlap the rendering of the next
    cudaEvent_t ev_sub[2];
    cudaStream_t aux_stream;         // code-NCO scan and tile prologues run beside the carrier chain / the sample kernels
    cudaEvent_t ev_fork, ev_fork2, ev_code2, ev_chain, ev_P[2], ev_F[2];
    cudaEvent_t ev[TIMING_RING][5];  // per recorded step: begin, scans done (= render start), render done,
                                     // and around the first k_synth_line launch of the step
    int fixed_epochs;                // epochs covered by that launch
    int ev_count;                    // steps recorded since gpsiq_timing_begin
    gpsiq_chan_desc* d_desc;
    int2* d_lut;          // [E][C][512]
    int32_t* d_lutp;      // [E][C][512] packed (Q << 16) + I
    int* d_flags;         // [2][E]: amplitude sum per epoch, step-contract flag per epoch
    double* d_bias_rate;  // [C] measured residual of the closed-form epoch advance (cycles per epoch), see k_bias_update
    double* d_carr_start; // [C] exact phases at the start of the batch being chained
    int chain_keeps_estimate;  // GPSIQ_OPT_CHAIN_KEEPS_ESTIMATE
    Handoff handoff;           // what the next enqueue_chain fuses into its kernel (gpsiq_chain_handoff_device); cleared after use
    int free_running;          // GPSIQ_OPT_FREE_RUNNING_ESTIMATE
    double* d_est_meas;        // [C] latest open-loop measurement of the estimate's accumulated error
    double* d_est_ccum;        // [C] total correction the running estimate has received
    int slice_spec;            // level 5 (slice-level speculation: one head scan per batch on the chain's critical path); GPSIQ_SLICE_SPEC=0 turns it off
    int render_after_next_chain;  // GPSIQ_OPT_RENDER_AFTER_NEXT_CHAIN
    // SM-free carrier hand-off between the GPUs of one node (gpsiq_mailbox_*)
    unsigned char* d_mbox;      // own mailbox: [2][MBOX_SLOT] state slots, sequence flag at MBOX_FLAG
    unsigned char* d_mbox_peer; // the next rank's mailbox (CUDA IPC mapping)
    void* fn_write64;           // cuStreamWriteValue64 / cuStreamWaitValue64
    void* fn_wait64;
    int mbox_flush;             // the device can flush remote writes after a wait
    int line_grid_cap;    // GPSIQ_OPT_LINE_GRID_CAP: most CTAs of one k_synth_line launch (0: one CTA per unit)
    int use_line;         // k_synth_line (the production kernel) is eligible for this configuration
    int8_t* d_chips4;     // [33][4][LN_VS] +-1: chip/NAV sign tables in 4 polarity variants, extended past chip 1022
    ulonglong2* d_anch[NSETS];   // [E][ntiles][C] tile anchors {F, G} (one buffer per scan set)
    LineEpoch* d_lrecs;   // [E][C] per-epoch line records of the batch being rendered (k_line_anchor -> k_line_check)
    uint32_t* d_elist;    // (epoch, slot) pairs k_line_check could not clear
    uint32_t* d_hazlist;  // (tile, slot) pairs k_line_refine could not clear
    int haz_cap, patch_cap;
    int* d_line_counters; // [0] listed hazards, [1] patches, [2] flagged chunks (per batch)
    unsigned long long* d_line_totals;  // the same, accumulated over the context's life
    LinePatch* d_patches;
    struct { const gpsiq_chan_desc* desc; int16_t* iq; int ne, set, e0; } last_ln;  // last k_synth_line launch
    double* d_code_ck;    // [E][ntiles][C]
    int* d_wrap_ck;       // [E][ntiles][C]
    double* d_carr_ck;    // [5][E][ntiles][C] planes: 0,1 chunk speculation (parity variants), 2,3 stitched epoch-level
                          // trajectory P (variants), 4 exact (chain heads / fallbacks; INT32 mode: uint32 phase as double)
    double* d_drift;      // [E][C] estimate aids only: closed-form phase advance of each epoch incl. predicted
                          // rounding drift (eadv), then [E][C] re-seed phase or -1 (ereset), then [E][C] the
                          // estimated phase at the start of each epoch (est_epoch)
    CarrSpec* d_spec;     // [E][C][J][2] chunk-level speculation results
    CarrSpec* d_specE;    // [E][C][2]    epoch-level (stitched) results
    ChunkInfo* d_cinfo;   // [E][C][2][J]
    int G, J;             // chunk length in tiles, chunks per epoch
    CarrInfo* d_info;     // [3][E][C]: epoch results of the group-chain variants 0, 1 and of the exact fallback chain
    CarrSpec* d_specG;    // [groups][C][2] group-level speculation results
    GroupInfo* d_ginfo;   // [groups][C]    final chain results
    double* d_traceG;     // [2][E][C]      post-epoch phases of the group chains
    CarrSpec* d_specS;    // [C][2]         slice-level speculation results (level 5)
    double* d_startS;     // [2][groups][C] phase of the slice-level trajectories at every group start
    SliceRes* d_sres;     // [C]            how k_carr_final passed the batch (translated / serial)
    TieEvent* d_tieG;     // [groups][C][2] first tie-wrap of the group-level trajectories (nco_scan.cuh: TieEvent)
    TieEvent* d_tieS;     // [C][2]         ... of the slice-level trajectories (pos = group index)
    double* d_start0;     // [C]            exact phases at the batch's first sample
    int* d_fallbacks;     // epochs that fell back to the serial carrier scan (diagnostic counter)
    unsigned long long* d_slice_stats;  // [2] (slot, batch) chains passed by the slice-level translation / chained serially
    size_t ck_plane;      // elements per plane
    double* d_carr_state; // [C]  exact carrier phase per slot after the last chained epoch
    double* d_est_state;  // [C]  ESTIMATED phase at the start of the next batch to speculate (never part of a result)
    double* d_adv;        // [2C] this batch's closed-form phase advance per slot + "re-seeded" flags
    double* d_carr_trace; // [E][C]
    uint32_t* d_ustart;   // [E][C] integer-carrier mode: the uint32 phase at the first sample of every epoch
    uint32_t* d_ca;       // [33][CA_WORDS]
    int16_t* d_iq;        // [E][N][2]
    int16_t* d_iq2;       // second output buffer for the host streaming pair (allocated on first use)
    gpsiq_chan_desc* h_stage[NSETS];
    double* d_adv_prev;           // [2C] multi-device streams: the previous batch's closed-form advance, copied from its device
    long long seq;                // batches begun so far
    cudaEvent_t ev_final;         // the last exact chain enqueued (on any stream) has run
    int fetch_count;              // host fetches so far (alternates the two device output buffers)
    int render_waits_spec;        // see enqueue_render  // pinned staging of submitted host descriptors
    unsigned long long* d_sums;
    int* d_err;
    int last_epochs;
    int64_t launches;
    int trace_on, trace_n;
    TraceRec* trace;
    char err[256];
};

static void trace_mark(gpsiq_ctx* ctx, cudaStream_t st, const char* label) {
    if (!ctx->trace_on) return;
    if (ctx->trace_n >= TRACE_MAX) ctx->trace_n = 0;  // ring
    TraceRec& r = ctx->trace[ctx->trace_n];
    if (!r.ev) cudaEventCreate(&r.ev);
    r.label = label;
    r.stream_id = -1;
    for (int i = 0; i < NSETS; i++) if (st == ctx->sets[i].stream) r.stream_id = 10 + i;
    if (r.stream_id < 0) r.stream_id = (st == ctx->scan_stream) ? 1 : (st == ctx->aux2_stream) ? 2 : (st == ctx->aux_stream) ? 3 : (st == ctx->copy_stream) ? 4 : -1;
    if (r.stream_id < 0) {  // a caller's stream: numbered 20, 21, ... in order of first appearance
        static cudaStream_t seen[16];
        static int nseen = 0;
        int k = 0;
        while (k < nseen && seen[k] != st) k++;
        if (k == nseen && nseen < 16) seen[nseen++] = st;
        r.stream_id = 20 + k;
    }
    {
        struct timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        r.host_ms = (double) ts.tv_sec * 1e3 + (double) ts.tv_nsec * 1e-6;
    }
    cudaEventRecord(r.ev, st);
    ctx->trace_n++;
}

static char g_err[256];


static int fail(gpsiq_ctx* ctx, int code, const char* what, cudaError_t ce) {
    char* dst = ctx ? ctx->err : g_err;
    if (ce != cudaSuccess)
        snprintf(dst, 256, "%s: %s", what, cudaGetErrorString(ce));
    else
        snprintf(dst, 256, "%s", what);
    return code;
}

#define CU(call)                                                             \
    do {                                                                     \
        cudaError_t ce_ = (call);                                            \
        if (ce_ != cudaSuccess) return fail(ctx, GPSIQ_ERR_CUDA, #call, ce_); \
    } while (0)

// ---------------------------------------------------------------------------
// k_prepare: amplitude LUT per (epoch, slot)
// ---------------------------------------------------------------------------
__global__ void k_prepare(const gpsiq_chan_desc* __restrict__ desc, int2* __restrict__ lut,
                          int32_t* __restrict__ lutp, double* __restrict__ drift,
                          int* __restrict__ amp_sum, int* __restrict__ step_flag, int C, int N, int carrier_mode,
                          const double* __restrict__ bias_rate, double* __restrict__ eadv_frac, int* __restrict__ err) {
    const int ec = blockIdx.x;
    const gpsiq_chan_desc d = desc[ec];
    int2* out = lut + (size_t) ec * 512;
    if (d.prn <= 0) {
        if (threadIdx.x == 0) { drift[ec] = 0.0; drift[(size_t) gridDim.x + ec] = -1.0; eadv_frac[ec] = 0.0; }
        return;
    }
    if (threadIdx.x == 32 && carrier_mode == GPSIQ_CARRIER_FLOAT) {
        const double dest = carr_drift_estimate(d.carr_step, step_info(d.carr_step), N);
        drift[ec] = fma((double) N, d.carr_step, dest);   // eadv: whole-epoch advance in cycles (chunk interpolation)
        // The same advance modulo one cycle, to full precision: N*step is ~300 cycles, so its double carries only
        // ~5e-14 of absolute precision -- more than the per-epoch corrections that matter here.  Split the product
        // exactly (fma), take the fraction of the big part, and add the small terms: the predicted rounding drift
        // and the slot's measured residual per epoch (k_bias_update).
        const double p = __dmul_rn((double) N, d.carr_step);
        const double pe = fma((double) N, d.carr_step, -p);
        eadv_frac[ec] = (p - floor(p)) + ((pe + dest) - bias_rate[ec % C]);
        drift[(size_t) gridDim.x + ec] = (d.flags & GPSIQ_FLAG_RESET_CARRIER) ? d.carr_phase0 : -1.0;  // ereset
    }
    // (the descriptor's NAV window is 64 bits: an epoch may not reach past it -- the kernels index it modulo 64 and the
    // reference's own bit fetch, plutogpssim.c:2732, has no such limit)
    const bool nav_window_ok = ((double) (d.ms0 % 20) + (double) N * d.code_step / 1023.0 + 1.0) / 20.0 < 64.0;
    if (d.prn > 32 || !(d.code_phase0 >= 0.0 && d.code_phase0 < 1023.0) || !(d.code_step > 0.0 && d.code_step < 1023.0) ||
        !nav_window_ok) {
        if (threadIdx.x == 0) atomicExch(err, 1 + ec);
        return;
    }
    for (int k = threadIdx.x; k < 512; k += blockDim.x) {
        // (int)(+-1 * table * gain) == +-(int)(table * gain): int->double is exact,
        // one rounding in the product, truncation toward zero (plutogpssim.c:2701-2702)
        int ip = __double2int_rz(__dmul_rn((double) c_cos512[k], d.gain));
        int qp = __double2int_rz(__dmul_rn((double) c_sin512[k], d.gain));
        out[k] = make_int2(ip, qp);
        lutp[(size_t) ec * 512 + k] = qp * 65536 + ip;  // exact while the per-sample sums fit int16 (amp_sum)
    }
    if (threadIdx.x == 64) {
        // |entry| <= 512*|gain| + 1: bound on this slot's share of |sum I|, |sum Q|
        const double a = fabs(d.gain) * 512.0 + 1.0;
        atomicAdd(&amp_sum[ec / C], a < 40000.0 ? (int) a : 40000);
        // contract of k_synth_line: <= 4 carrier cycles and <= 1 code-period wrap per 1024-sample tile
        // (integer carrier: the step is in counts of 2^-25 cycle)
        const double cmax = (carrier_mode == GPSIQ_CARRIER_FLOAT) ? 0x1p-8 : 0x1p17;
        if (!(fabs(d.carr_step) <= cmax) || !(d.code_step <= 0.5)) atomicOr(&step_flag[ec / C], 1);
    }
}

// ---------------------------------------------------------------------------
// k_scan_code: exact code phase + wrap count at every tile start
// ---------------------------------------------------------------------------
// Needed only by k_synth_lanes (the literal-recurrence kernel): k_synth_line's code anchors are closed form.
// flags != NULL: only the epochs k_synth_lanes will render (amplitude / step contract, list overflows).
__global__ void k_scan_code(const gpsiq_chan_desc* __restrict__ desc,
                            double* __restrict__ code_ck, int* __restrict__ wrap_ck, const int* __restrict__ amp_sum,
                            const int* __restrict__ step_flag, int EC, int C, int N, int T, int ntiles) {
    // one chain per thread; the lanes of a warp hold the same slot for 32 consecutive epochs (same
    // satellite, nearly the same code rate), so their segment walks stay mostly convergent
    const int chain = blockIdx.x * blockDim.x + threadIdx.x;
    if (chain >= EC) return;
    const int E = EC / C;
    const int e = chain % E, c = chain / E;
    const int ec = e * C + c;
    if (amp_sum && !(amp_sum[e] > 32767 || step_flag[e])) return;
    const gpsiq_chan_desc d = desc[ec];
    if (d.prn <= 0) return;
    double x = d.code_phase0;
    int wraps = 0;
    const StepInfo tab = step_info(d.code_step);
    for (int t = 0; t < ntiles; t++) {
        const size_t o = ((size_t) e * ntiles + t) * C + c;
        code_ck[o] = x;
        wrap_ck[o] = wraps;
        const int len = min(T, N - t * T);
        nco_advance<NCO_CODE>(x, d.code_step, tab, len, wraps);
    }
}

// ---------------------------------------------------------------------------
// k_scan_carrier: exact carrier phase at every tile start, chained over epochs
// ---------------------------------------------------------------------------
__global__ void k_scan_carrier(const gpsiq_chan_desc* __restrict__ desc,
                               double* __restrict__ carr_ck,
                               double* __restrict__ carr_state, double* __restrict__ carr_trace,
                               CarrInfo* __restrict__ info, GroupInfo* __restrict__ ginfo, int GP, int E, int C, int N,
                               int T, int ntiles, int carrier_mode) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // one chain per warp, lane 0
    if (c >= C || (threadIdx.x & 31)) return;
    for (int g = 0; g * GP < E; g++) {  // every group: "chained exactly", all tiles in the exact plane
        GroupInfo gi; gi.delta = 0.0; gi.pos = 0x7fffffff; gi.variant = 0;
        ginfo[(size_t) g * C + c] = gi;
    }
    double x = carr_state[c];
    int dummy = 0;
    for (int e = 0; e < E; e++) {
        const gpsiq_chan_desc d = desc[(size_t) e * C + c];
        {
            CarrInfo inf; inf.delta = 0.0; inf.n1 = N; inf.variant = 0;  // every tile reads the exact plane
            info[(size_t) e * C + c] = inf;
        }
        if (d.prn <= 0) {
            carr_trace[(size_t) e * C + c] = x;
            continue;
        }
        if (d.flags & GPSIQ_FLAG_RESET_CARRIER) x = d.carr_phase0;
        const StepInfo tab = step_info(d.carr_step);
        for (int t = 0; t < ntiles; t++) {
            carr_ck[((size_t) e * ntiles + t) * C + c] = x;  // (carr_ck = the exact plane)
            const int len = min(T, N - t * T);
            nco_advance<NCO_CARRIER>(x, d.carr_step, tab, len, dummy);
        }
        carr_trace[(size_t) e * C + c] = x;
    }
    carr_state[c] = x;
}

// ---------------------------------------------------------------------------
// k_int_carrier: the integer carrier NCO (GPSIQ_CARRIER_INT32; plutogpssim.c:1966-1967, 2675, 2699, 2748).
// phase += step per sample modulo 2^32 is a closed form, so the whole "chain" is one prefix sum over the
// batch's epochs per slot: ustart[e][c] = phase at the first sample of epoch e (tile starts follow as
// ustart + step * n, evaluated where they are needed), trace = phase after every epoch.
// One warp per slot: the lanes stage the slot's column of steps / re-seeds in shared memory, lane 0 runs
// the short recurrence.  mode 0 (chain): from carr_state, which it advances.  mode 1 (advance, used by the
// prepare phase of time-sliced runs): from zero, result to adv[c] with adv[C + c] = 1 if a descriptor re-seeded
// the slot (adv[c] is then an absolute phase) -- the closed-form hand-off of SURVEY 8e: no ring, a prefix.
// ---------------------------------------------------------------------------
#define INTC_MAX_E 2048
__global__ void __launch_bounds__(32)
k_int_carrier(const gpsiq_chan_desc* __restrict__ desc, uint32_t* __restrict__ ustart, double* __restrict__ carr_state,
              double* __restrict__ carr_trace, double* __restrict__ adv, int mode, int E, int C, int N) {
    __shared__ uint32_t s_step[INTC_MAX_E], s_seed[INTC_MAX_E];
    __shared__ uint8_t s_kind[INTC_MAX_E];   // 0 inactive, 1 active, 2 active + re-seeded
    const int c = blockIdx.x, lane = threadIdx.x;
    if (c >= C) return;
    uint32_t u = (mode == 0) ? (uint32_t) carr_state[c] : 0u;
    bool seeded = false;
    for (int e0 = 0; e0 < E; e0 += INTC_MAX_E) {
        const int n = min(INTC_MAX_E, E - e0);
        for (int i = lane; i < n; i += 32) {
            const gpsiq_chan_desc& d = desc[(size_t) (e0 + i) * C + c];
            const int prn = d.prn;
            s_kind[i] = prn <= 0 ? 0 : ((d.flags & GPSIQ_FLAG_RESET_CARRIER) ? 2 : 1);
            s_step[i] = prn > 0 ? (uint32_t) (int32_t) d.carr_step * (uint32_t) N : 0u;
            s_seed[i] = prn > 0 ? (uint32_t) d.carr_phase0 : 0u;
        }
        __syncwarp();
        if (lane == 0)
            for (int i = 0; i < n; i++) {
                if (s_kind[i] == 2) { u = s_seed[i]; seeded = true; }
                s_seed[i] = u;                       // (becomes the ustart column)
                u += s_step[i];
                s_step[i] = u;                       // (becomes the trace column)
            }
        __syncwarp();
        if (mode == 0)
            for (int i = lane; i < n; i += 32) {
                ustart[(size_t) (e0 + i) * C + c] = s_seed[i];
                carr_trace[(size_t) (e0 + i) * C + c] = (double) s_step[i];
            }
        __syncwarp();
    }
    if (lane == 0) {
        if (mode == 0) carr_state[c] = (double) u;
        else { adv[c] = (double) u; adv[C + c] = seeded ? 1.0 : 0.0; }
    }
}

// carr_state <- fold(carr_state, adv) in the integer carrier's arithmetic (exact): skip over a slice synthesized
// elsewhere.  adv as written by k_int_carrier mode 1.
__global__ void k_int_fold(double* __restrict__ state, const double* __restrict__ adv, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const uint32_t a = (uint32_t) adv[c];
    state[c] = (double) ((adv[C + c] != 0.0) ? a : (uint32_t) state[c] + a);
}


// ---------------------------------------------------------------------------
// Parallel carrier scan (FLOAT mode), see nco_scan.cuh "speculate -> translate -> verify".
// k_carr_speculate: one chain per (epoch, slot, parity variant), all independent.
// k_carr_chain    : one chain per slot, serial over epochs, O(one carrier cycle) per epoch.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double est_advance_dev(double x, double d, int N) {
    double t = fma((double) N, d, x);
    t -= floor(t);
    return (t >= 0.0 && t < 1.0) ? t : 0.0;
}

__device__ __forceinline__ double frac01(double t) {
    t -= floor(t);
    return (t >= 0.0 && t < 1.0) ? t : 0.0;
}

// Closed-form (estimated) effect of one batch on the carrier phase of every slot:
// adv[c] = sum over the batch's epochs of N*step + predicted rounding drift, and
// adv[C+c] = 1 if a descriptor re-seeded the slot (then adv[c] is an absolute phase).
// Feeds only the start-phase ESTIMATES of later speculative scans, on this GPU or
// -- for time-sliced multi-GPU runs -- on the ranks that own later slices.
// One warp per slot: the lanes stage the slot's column of the two [E][C] arrays in shared memory
// (a lane-strided, latency-overlapped load), then lane 0 runs the short serial recurrence from there.
#define EST_MAX_E 2048
__device__ __forceinline__ void stage_column(double* s_a, double* s_r, const double* __restrict__ eadv,
                                             const double* __restrict__ ereset, int e0, int n, int c, int C, int lane) {
    for (int i = lane; i < n; i += 32) {
        s_a[i] = eadv[(size_t) (e0 + i) * C + c];
        s_r[i] = ereset[(size_t) (e0 + i) * C + c];
    }
    __syncwarp();
}

__global__ void __launch_bounds__(32)
k_slice_advance(const double* __restrict__ eadv, const double* __restrict__ ereset, double* __restrict__ adv, int E, int C) {
    __shared__ double s_a[EST_MAX_E], s_r[EST_MAX_E];
    const int c = blockIdx.x, lane = threadIdx.x;
    if (c >= C) return;
    double x = 0.0, abs_flag = 0.0;
    for (int e0 = 0; e0 < E; e0 += EST_MAX_E) {
        const int n = min(EST_MAX_E, E - e0);
        stage_column(s_a, s_r, eadv, ereset, e0, n, c, C, lane);
        if (lane == 0)
            for (int i = 0; i < n; i++) {
                const double r = s_r[i];
                if (r >= 0.0) { x = r; abs_flag = 1.0; }
                x = frac01(x + s_a[i]);
            }
        __syncwarp();
    }
    if (lane == 0) { adv[c] = x; adv[C + c] = abs_flag; }
}

// After the exact chain of a batch: how far was the closed-form advance off?  adv[c] is the predicted advance of the
// slot over the batch (k_slice_advance), start/end its exact phases before/after the chain.  The per-epoch residual
// is integrated into rate[c], which k_prepare subtracts from later predictions: the predictor's systematic error
// (~1e-14 cycles per epoch) otherwise grows along a batch until group-level speculations stop fitting.
// Estimates only: never part of a result.
// OPEN LOOP: what is measured is the residual of the UNCORRECTED closed form (the correction the batch was prepared with,
// rate_used, is added back), and rate[c] is a filter of that measurement -- not an integrator of the corrected error.
// Batches are prepared one or two batches ahead of their chain (pipelined submits, the pipelined time-slice runner);
// an integrator with that much delay in its loop oscillates (gain 0.7, two batches of delay: unstable), a filter of an
// open-loop measurement does not care.
__global__ void k_bias_update(const double* __restrict__ adv, const double* __restrict__ start,
                              const double* __restrict__ end, double* __restrict__ rate,
                              const double* __restrict__ rate_used, int n_epochs, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C || adv[C + c] != 0.0) return;  // re-seeded inside the batch: adv is an absolute phase
    double err = adv[c] - (end[c] - start[c]);
    err -= rint(err);
    if (fabs(err) < 1e-8) {
        const double raw = err / (double) n_epochs + rate_used[c];   // residual per epoch of the closed form itself
        rate[c] += 0.9 * (raw - rate[c]);   // (the measurement is exact arithmetic over the whole batch: little to smooth)
    }
}

// est = fold(est, adv): est <- adv (absolute) or frac(est + adv)
__global__ void k_est_fold(double* __restrict__ est, const double* __restrict__ adv, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    est[c] = frac01((adv[C + c] != 0.0) ? adv[c] : est[c] + adv[c]);
}

// est <- est - gain * (est_old - exact_old): feeds a measured estimate error of an earlier slice back
// (time-sliced runs, where the estimate is never re-anchored on the exact phase directly)
__global__ void k_est_correct(double* __restrict__ est, const double* __restrict__ exact_old,
                              const double* __restrict__ est_old, double gain, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double diff = est_old[c] - exact_old[c];
    diff -= rint(diff);
    est[c] = frac01(est[c] - gain * diff);
}

// ---- free-running start-phase estimates (GPSIQ_OPT_FREE_RUNNING_ESTIMATE; time-sliced runs) ----------------------
// The speculation of a rank's NEXT slice must not wait for the exact chain of this one (the chain is a hop of the
// inter-GPU ring).  So the estimate runs on its own: after a slice has been speculated it becomes the END of that
// slice's own slice-level speculative trajectory (k_est_from_slice: an exact advance from the estimated start), the
// caller folds in the closed-form advances of the slices other GPUs own, and the error this accumulates is taken out
// OPEN LOOP: every exact chain measures eps = (estimate its slice started from) - (exact start phase), adds back the
// total correction that estimate had already received (ccum_used) and publishes M = eps + ccum_used -- a measurement of
// the raw accumulated error that does not depend on what was corrected when; the next speculation subtracts
// (M - ccum) and sets ccum = M.  No integrator, so the two-slice delay between measurement and use cannot oscillate.
// Estimates only: never part of a result.
__global__ void k_est_open_loop(double* __restrict__ est, const double* __restrict__ meas, double* __restrict__ ccum,
                                double* __restrict__ ccum_used, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double m = meas[c];
    est[c] = frac01(est[c] - (m - ccum[c]));
    ccum[c] = m;
    ccum_used[c] = m;
}

// est <- the end of the slice-level speculative trajectory of the batch just speculated (variant 0; a parity shift of
// 2^-53 does not matter to an estimate), or -- if that chain did not run to the end -- its start advanced in closed form
__global__ void k_est_from_slice(double* __restrict__ est, const CarrSpec* __restrict__ specS,
                                 const double* __restrict__ startS, const double* __restrict__ adv, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const CarrSpec s0 = specS[(size_t) c * 2];
    if (s0.pad & 2) est[c] = s0.xend;                                  // the chain ran through every group
    else est[c] = frac01((adv[C + c] != 0.0) ? adv[c] : startS[c] + adv[c]);
}

// Estimated phase at the start of every epoch of the batch, from the context's batch-start estimate.
__global__ void __launch_bounds__(32)
k_epoch_estimates(const double* __restrict__ eadv, const double* __restrict__ ereset,
                  const double* __restrict__ est_state, double* __restrict__ est_epoch, int E, int C) {
    __shared__ double s_a[EST_MAX_E], s_r[EST_MAX_E];
    const int c = blockIdx.x, lane = threadIdx.x;
    if (c >= C) return;
    double x = est_state[c];
    for (int e0 = 0; e0 < E; e0 += EST_MAX_E) {
        const int n = min(EST_MAX_E, E - e0);
        stage_column(s_a, s_r, eadv, ereset, e0, n, c, C, lane);
        if (lane == 0)
            for (int i = 0; i < n; i++) {
                const double r = s_r[i];
                if (r >= 0.0) x = r;  // re-seeded: the start phase of this epoch is known exactly
                const double a = s_a[i];
                s_a[i] = x;           // (becomes the output column)
                x = frac01(x + a);
            }
        __syncwarp();
        for (int i = lane; i < n; i += 32) est_epoch[(size_t) (e0 + i) * C + c] = s_a[i];
        __syncwarp();
    }
}

// One chain per (epoch, slot, chunk): speculative scan of the chunk's tiles, both parity variants (the second one is
// derived from the first, nco_scan.cuh: spec_scan_chunk).
__global__ void k_carr_speculate(const gpsiq_chan_desc* __restrict__ desc,
                                 const double* __restrict__ eadv, const double* __restrict__ est_epoch,
                                 double* __restrict__ carr_ck, size_t ck_plane, CarrSpec* __restrict__ spec, int E,
                                 int C, int N, int T, int ntiles, int G, int J) {
    // one chain per THREAD; the lanes of a warp hold the same (slot, chunk) for 32 consecutive
    // epochs: same satellite, nearly the same Doppler, so their segment walks stay mostly convergent
    const int chain0 = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = chain0 < E * C * J;
    const int chain = in_range ? chain0 : 0;
    const int e = chain % E;
    const int rest = chain / E;
    const int j = rest % J, c = rest / J;
    const int ec = e * C + c;
    const gpsiq_chan_desc d = desc[ec];
    CarrSpec out0, out1;
    out0.margin = -1.0; out0.n1 = -1; out0.xw1 = 0.0; out0.xend = 0.0; out0.pad = 0;
    out1 = out0;
    const bool run = in_range && d.prn > 0 && carr_step_speculable(d.carr_step);
    const unsigned mask = __ballot_sync(0xffffffffu, run);  // the lanes that walk together (lockstep loops, nco_scan.cuh)
    if (run) {
        const int t0 = j * G, t1 = min(t0 + G, ntiles);
        // estimated phase at the chunk's first sample (chunk 0: the epoch estimate itself)
        double x = est_epoch[ec];
        if (j > 0) x = frac01(x + eadv[ec] * ((double) (t0 * T) / (double) N));
        double* ck0 = carr_ck + (size_t) e * ntiles * C + c;
        spec_scan_chunk(x, d.carr_step, step_info(d.carr_step), N, T, t0, t1, ck0, ck0 + ck_plane, (size_t) C, out0, out1, mask);
    }
    if (in_range) {
        spec[((size_t) ec * J + j) * 2] = out0;
        spec[((size_t) ec * J + j) * 2 + 1] = out1;
    }
}

#define SPEC_MAX_CHUNKS 16  // most chunks per epoch (level 1)
// One chain per (epoch, slot, epoch-level variant), one per THREAD (lanes = 32 consecutive epochs of one satellite, like
// the chunk speculation): stitch the chunk runs into the epoch-level trajectory P.
__global__ void k_carr_stitch(const gpsiq_chan_desc* __restrict__ desc,
                              const double* __restrict__ est_epoch, const CarrSpec* __restrict__ spec,
                              double* __restrict__ carr_ck, size_t ck_plane, ChunkInfo* __restrict__ cinfo,
                              CarrSpec* __restrict__ specE, int E, int C, int N, int T, int ntiles, int G, int J) {
    const int chain = blockIdx.x * blockDim.x + threadIdx.x;
    if (chain >= E * C * 2) return;
    const int e = chain % E;
    const int rest = chain / E;
    const int V = rest & 1, c = rest >> 1;
    const int ec = e * C + c;
    const gpsiq_chan_desc d = desc[ec];
    CarrSpec out;
    out.margin = -1.0; out.n1 = -1; out.xw1 = 0.0; out.xend = 0.0; out.pad = 0;
    if (d.prn > 0 && !(V == 1 && d.carr_step >= 0.0) && carr_step_speculable(d.carr_step))
        stitch_epoch(est_epoch[ec], d.carr_step, step_info(d.carr_step), N, T, G, V, spec + (size_t) ec * J * 2,
                     carr_ck + (size_t) (2 + V) * ck_plane + (size_t) e * ntiles * C + c, (size_t) C,
                     cinfo + ((size_t) ec * 2 + V) * J, out);
    specE[(size_t) ec * 2 + V] = out;
}

// Stage the per-epoch inputs of one group for one slot into shared memory (one lane per epoch).
__device__ __forceinline__ void stage_group(GroupEpoch* ge, const gpsiq_chan_desc* __restrict__ desc,
                                            const CarrSpec* __restrict__ specE,
                                            int first, int count, int c, int C, int lane) {
    for (int i = lane; i < count; i += 32) {
        const size_t ec = (size_t) (first + i) * C + c;
        const gpsiq_chan_desc d = desc[ec];
        GroupEpoch& g = ge[i];
        g.d = d.carr_step;
        g.phase0 = d.carr_phase0;
        g.active = d.prn > 0;
        g.reset = (d.flags & GPSIQ_FLAG_RESET_CARRIER) != 0;
        g.s0 = specE[ec * 2];
        g.s1 = specE[ec * 2 + 1];
    }
    __syncwarp();
}

#define GROUP_EPOCHS 64  // epochs per group (level 3): the final, serial chain does one head scan per group

// Level 3: one chain per (group, slot, variant): the group's epochs chained from the ESTIMATED group start.
// One warp per block: 4.6 KB of shared memory, so that the blocks fit beside two resident k_synth_line CTAs.
__global__ void __launch_bounds__(32)
k_carr_group(const gpsiq_chan_desc* __restrict__ desc,
             const CarrSpec* __restrict__ specE, const double* __restrict__ est_epoch, double* __restrict__ carr_ck,
             size_t ck_plane, CarrInfo* __restrict__ infoG, size_t info_plane, double* __restrict__ traceG,
             CarrSpec* __restrict__ specG, TieEvent* __restrict__ tieG, int* __restrict__ fallbacks, int E, int C, int N,
             int T, int ntiles) {
    __shared__ GroupEpoch s_ge[GROUP_EPOCHS];
    const int lane = threadIdx.x;
    const int ngroups = (E + GROUP_EPOCHS - 1) / GROUP_EPOCHS;
    const int chain = blockIdx.x;
    if (chain >= ngroups * C * 2) return;
    const int V = chain & 1, gc = chain >> 1;
    const int g = gc / C, c = gc - g * C;
    const int first = g * GROUP_EPOCHS, count = min(GROUP_EPOCHS, E - first);
    stage_group(s_ge, desc, specE, first, count, c, C, lane);
    if (lane) return;
    CarrSpec out;
    out.margin = -1.0; out.n1 = -1; out.xw1 = 0.0; out.xend = 0.0; out.pad = 0;
    TieEvent tie;
    tie.pos = -1; tie.k = 0;
    bool any_neg = false;
    for (int k = 0; k < count; k++) any_neg |= s_ge[k].active && s_ge[k].d < 0.0;
    if (V == 0 || any_neg) {
        int fb = 0;
        group_chain(est_epoch[(size_t) first * C + c], s_ge, count, N, T, V,
                    carr_ck + (size_t) (4 + V) * ck_plane + (size_t) first * ntiles * C + c, (size_t) C, (size_t) ntiles * C,
                    infoG + (size_t) V * info_plane + (size_t) first * C + c, (size_t) C,
                    traceG + (size_t) V * info_plane + (size_t) first * C + c, (size_t) C, out, tie, fb);
        if (fb && V == 0) atomicAdd(fallbacks, fb);
    }
    specG[(size_t) gc * 2 + V] = out;
    tieG[(size_t) gc * 2 + V] = tie;
}

// One group of the exact chain (level 4), executed by one warp: one head scan by lane 0 (fast path: only the group's
// first epoch is staged), the lanes translate the post-epoch phases of the other epochs in parallel.  x_start: the exact
// phase at the group's first sample (lane 0's value counts).  Returns the exact phase after the group (on lane 0).
__device__ __forceinline__ double final_group_warp(GroupEpoch* s_ge, double x_start, int g, const gpsiq_chan_desc* __restrict__ desc,
                                                   const CarrSpec* __restrict__ specE, const CarrSpec* __restrict__ specG,
                                                   const TieEvent* __restrict__ tieG, double* __restrict__ carr_ck, size_t ck_plane, CarrInfo* __restrict__ infoG,
                                                   size_t info_plane, const double* __restrict__ traceG,
                                                   double* __restrict__ carr_trace, GroupInfo* __restrict__ ginfo, int& fb,
                                                   int E, int C, int c, int N, int T, int ntiles, int lane) {
    const int first = g * GROUP_EPOCHS, count = min(GROUP_EPOCHS, E - first);
    const size_t o = (size_t) first * C + c;
    const CarrSpec sG0 = specG[((size_t) g * C + c) * 2], sG1 = specG[((size_t) g * C + c) * 2 + 1];
    const TieEvent tG0 = tieG[((size_t) g * C + c) * 2], tG1 = tieG[((size_t) g * C + c) * 2 + 1];
    double* ckX = carr_ck + 6 * ck_plane + (size_t) first * ntiles * C + c;
    // Fast path (almost always taken): the group's first epoch is active, wraps, and the group trajectory
    // fits from there.  Only that epoch's inputs are staged; the serial work is its head scan, and the
    // lanes translate the post-epoch phases of the other epochs in parallel.
    stage_group(s_ge, desc, specE, first, 1, c, C, lane);
    double x = x_start;
    GroupInfo gi;
    gi.delta = 0.0; gi.delta2 = 0.0; gi.pos = 0x7fffffff; gi.variant = 0; gi.tie_pos = 0x7fffffff; gi.pad = 0;
    int fb_try = 0;
    if (lane == 0)
        x = group_final(x_start, s_ge, 1, N, T, sG0, sG1, ckX, (size_t) C, (size_t) ntiles * C, infoG + 2 * info_plane + o,
                        (size_t) C, traceG + o, traceG + info_plane + o, carr_trace + o, (size_t) C, gi, fb_try, tG0, tG1);
    const int pos = __shfl_sync(0xffffffffu, gi.pos, 0);
    if (pos < N) {  // translated inside the first epoch
        const double diff = __shfl_sync(0xffffffffu, gi.delta, 0), diff2 = __shfl_sync(0xffffffffu, gi.delta2, 0);
        const int tie_pos = __shfl_sync(0xffffffffu, gi.tie_pos, 0);
        const int v = __shfl_sync(0xffffffffu, gi.variant, 0);
        const double* tg = traceG + (v ? info_plane : 0) + o;
        for (int e2 = 1 + lane; e2 < count; e2 += 32)   // (as group_final translates them: the shift after a tie event is diff2)
            carr_trace[o + (size_t) e2 * C] = add_rn(tg[(size_t) e2 * C], ((long long) tie_pos <= (long long) (e2 + 1) * N) ? diff2 : diff);
    } else {        // anything else: the general chain over the whole group, from the group's start state
        stage_group(s_ge, desc, specE, first, count, c, C, lane);
        if (lane == 0)
            x = group_final(x_start, s_ge, count, N, T, sG0, sG1, ckX, (size_t) C, (size_t) ntiles * C,
                            infoG + 2 * info_plane + o, (size_t) C, traceG + o, traceG + info_plane + o, carr_trace + o,
                            (size_t) C, gi, fb, tG0, tG1);
    }
    if (lane == 0) ginfo[(size_t) g * C + c] = gi;
    __syncwarp();
    return x;
}

// Result of the slice-level match of one slot (k_carr_final -> k_carr_final_groups)
struct SliceRes {     // how: 0 chained serially, 1 translated, 2 slot inactive in the whole batch
    double diff, diff2;   // translation of the slice-level trajectory; diff2: for the groups after group tie_g (TieEvent)
    int variant, how, tie_g, pad;
};
#define GPSIQ_DEVERR_SLICE 0x10000000   // error word: the parallel group chain disagreed with the slice-level translation

// Level 5, speculative (nco_scan.cuh: slice_chain_group): one chain per (slot, variant), serial over the batch's groups
// from the ESTIMATED batch start -- in the speculation phase, off the hand-off path of time-sliced runs.
// startS [2][ngroups][C]: the phase at which the chain enters every group; specS [C][2]: first wrap, end, margin.
__global__ void __launch_bounds__(32)
k_carr_slice(const gpsiq_chan_desc* __restrict__ desc, const CarrSpec* __restrict__ specE,
             const CarrSpec* __restrict__ specG, const TieEvent* __restrict__ tieG, const double* __restrict__ est_epoch,
             double* __restrict__ startS, CarrSpec* __restrict__ specS, TieEvent* __restrict__ tieS, int E, int C, int N) {
    __shared__ GroupEpoch s_ge[GROUP_EPOCHS];
    const int lane = threadIdx.x;
    const int V = blockIdx.x & 1, c = blockIdx.x >> 1;
    if (c >= C) return;
    const int ngroups = (E + GROUP_EPOCHS - 1) / GROUP_EPOCHS;
    double x = est_epoch[c];
    GroupTrack tr;
    tr.margin = 1.0; tr.xw1 = 0.0; tr.pos = -1; tr.usable = 1; tr.tie.pos = -1; tr.tie.k = 0;
    int any_active = 0;
    for (int g = 0; g < ngroups && tr.usable; g++) {
        const int first = g * GROUP_EPOCHS, count = min(GROUP_EPOCHS, E - first);
        const CarrSpec sG0 = specG[((size_t) g * C + c) * 2], sG1 = specG[((size_t) g * C + c) * 2 + 1];
        const TieEvent tG0 = tieG[((size_t) g * C + c) * 2], tG1 = tieG[((size_t) g * C + c) * 2 + 1];
        if (lane == 0) startS[((size_t) V * ngroups + g) * C + c] = x;
        stage_group(s_ge, desc, specE, first, 1, c, C, lane);       // fast path: the group's first epoch wraps
        int r = 0;
        if (lane == 0) r = slice_chain_group(x, s_ge, 1, N, sG0, sG1, tG0, tG1, tr, V, first, g, 0);
        r = __shfl_sync(0xffffffffu, r, 0);
        any_active |= __shfl_sync(0xffffffffu, s_ge[0].active, 0);
        if (r == 0 && count > 1) {                                   // it did not: the rest of the group, epoch by epoch
            __syncwarp();
            stage_group(s_ge, desc, specE, first, count, c, C, lane);
            for (int k = 1; k < count; k++) any_active |= s_ge[k].active;
            if (lane == 0) r = slice_chain_group(x, s_ge + 1, count - 1, N, sG0, sG1, tG0, tG1, tr, V, first, g, 1);
            r = __shfl_sync(0xffffffffu, r, 0);
        }
        if (r < 0) tr.usable = 0;
        __syncwarp();
    }
    if (lane == 0) {
        CarrSpec out;
        out.xw1 = tr.xw1; out.xend = x; out.n1 = tr.pos;
        out.pad = (any_active ? 0 : 1) | (tr.usable ? 2 : 0);   // bit 0: slot inactive in the whole batch; bit 1: the chain ran to the end
        out.margin = (tr.pos >= 0 && tr.usable) ? tr.margin : -1.0;
        specS[(size_t) c * 2 + V] = out;
        tieS[(size_t) c * 2 + V] = tr.tie;
    }
}

// Hand-off fused into the chain kernel (time-sliced multi-GPU runs, gpsiq_chain_handoff_device): the kernel itself
// waits for the previous slice's end phases in this GPU's mailbox (acquire load of a sequence flag the previous GPU
// writes over NVLink peer memory), chains, and stores its own end phases + sequence flag into the NEXT GPU's mailbox
// (peer stores, system-scope fence, release store by the last slot to finish).  One kernel per hop: no copy engine,
// no stream memory operation, nothing between the arrival of the phases and the head scan.
#define GPSIQ_DEVERR_HANDOFF 0x08000000  // error word: the previous slice's phases never arrived (bounded wait)

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Levels 5 + 4: the exact chain of a batch, one warp per slot.  With a usable slice-level speculation: ONE head scan
// (up to the batch's first wrap) + a translation, and k_carr_final_groups then chains the groups in parallel.  Otherwise
// serial over the groups: one head scan per group.  start_out / end_out: the exact phases before / after the batch.
__global__ void __launch_bounds__(32)
k_carr_final(const gpsiq_chan_desc* __restrict__ desc,
             const CarrSpec* __restrict__ specE, const CarrSpec* __restrict__ specG, double* __restrict__ carr_ck,
             size_t ck_plane, CarrInfo* __restrict__ infoG, size_t info_plane, const double* __restrict__ traceG,
             double* __restrict__ carr_state, double* __restrict__ carr_trace, GroupInfo* __restrict__ ginfo,
             int* __restrict__ fallbacks, const TieEvent* __restrict__ tieG, const CarrSpec* __restrict__ specS,
             const TieEvent* __restrict__ tieS, SliceRes* __restrict__ sres, unsigned long long* __restrict__ slice_stats,
             double* __restrict__ start_out, double* __restrict__ end_out, double* __restrict__ est_out,
             const Handoff h, const double* __restrict__ est_start, const double* __restrict__ ccum_used,
             double* __restrict__ est_meas, int E, int C, int N, int T, int ntiles) {
    __shared__ GroupEpoch s_ge[GROUP_EPOCHS];
    const int c = blockIdx.x, lane = threadIdx.x;
    if (c >= C) return;
    const int ngroups = (E + GROUP_EPOCHS - 1) / GROUP_EPOCHS;
    double x = 0.0;
    if (h.in_flag) {   // the start phases come from the previous slice's owner: wait for its message
        if (lane == 0) {
            const long long t0 = clock64();
            bool timed_out = false;
            while (ld_acquire_sys_u64(h.in_flag) < h.in_seq) {
                __nanosleep(100);
                if (clock64() - t0 > 20000000000LL) { timed_out = true; break; }   // ~10 s: give up loudly, never hang
            }
            if (timed_out) atomicOr(h.err, GPSIQ_DEVERR_HANDOFF);
            x = *(const volatile double*) (h.in_slot + c);
        }
        x = __shfl_sync(0xffffffffu, x, 0);
    } else {
        x = carr_state[c];
    }
    const double x_first = x;
    if (lane == 0) {
        start_out[c] = x;
        if (h.start_copy) h.start_copy[c] = x;
        if (est_meas) {   // free-running estimates: the raw accumulated error of the estimate this batch started from
            double eps = est_start[c] - x;
            eps -= rint(eps);
            if (fabs(eps) < 1e-8) est_meas[c] = eps + ccum_used[c];
        }
    }
    SliceRes res;
    res.diff = 0.0; res.diff2 = 0.0; res.variant = 0; res.how = 0; res.tie_g = 0x7fffffff; res.pad = 0;
    if (specS) {
        const CarrSpec sS0 = specS[(size_t) c * 2], sS1 = specS[(size_t) c * 2 + 1];
        const TieEvent tS0 = tieS[(size_t) c * 2], tS1 = tieS[(size_t) c * 2 + 1];
        if (sS0.pad & 1) {
            res.how = 2;                                            // nothing to chain: the phase passes through
        } else if (sS0.margin > 0.0) {
            const int count = min(GROUP_EPOCHS, E);
            stage_group(s_ge, desc, specE, 0, 1, c, C, lane);
            int r = 0;
            if (lane == 0) r = slice_verify(x, s_ge, 1, N, sS0, sS1, tS0, tS1, res.variant, res.diff, res.diff2, 0);
            r = __shfl_sync(0xffffffffu, r, 0);
            if (r == 0 && count > 1) {                              // no wrap in the first epoch: on through the first group
                __syncwarp();
                stage_group(s_ge, desc, specE, 0, count, c, C, lane);
                if (lane == 0) r = slice_verify(x, s_ge + 1, count - 1, N, sS0, sS1, tS0, tS1, res.variant, res.diff, res.diff2, 1);
                r = __shfl_sync(0xffffffffu, r, 0);
            }
            if (r == 1) { res.how = 1; const TieEvent& tv = res.variant ? tS1 : tS0; if (tv.pos >= 0) res.tie_g = tv.pos; }
            else x = x_first;                                 // (slice_verify leaves x advanced over the head on failure)
            __syncwarp();
        }
    }
    int fb = 0;
    if (res.how == 0) {
        for (int g = 0; g < ngroups; g++)
            x = final_group_warp(s_ge, x, g, desc, specE, specG, tieG, carr_ck, ck_plane, infoG, info_plane, traceG, carr_trace,
                                 ginfo, fb, E, C, c, N, T, ntiles, lane);
    }
    if (lane == 0) {
        if (sres) sres[c] = res;
        if (slice_stats && res.how != 2) atomicAdd(slice_stats + (res.how == 1 ? 0 : 1), 1ULL);
        if (h.out_slot) {   // end phases -> the next GPU's mailbox; the last slot to get here publishes the sequence number
            *(volatile double*) (h.out_slot + c) = x;
            __threadfence_system();
            if (atomicAdd(h.counter, 1u) == (unsigned) C - 1u) {
                __threadfence_system();   // (the other slots' stores, ordered before their increments, before the flag)
                *h.counter = 0u;
                st_release_sys_u64(h.out_flag, h.out_seq);
            }
        }
        carr_state[c] = x;
        end_out[c] = x;
        if (est_out) est_out[c] = x;
        if (fb) atomicAdd(fallbacks, fb);
    }
}

// Level 4 for the slots k_carr_final passed by translation: every group's exact start phase is known (the slice-level
// trajectory's phase at the group start + the translation), so the groups are chained in PARALLEL, one warp per
// (group, slot) -- after the hand-off of the end phases, not in front of it.  Writes exactly what the serial chain
// writes (exact plane, per-epoch results, post-epoch phases, GroupInfo) and checks that every group ends where the
// next one starts: a disagreement is an internal error, reported through the context's error word.
__global__ void __launch_bounds__(32)
k_carr_final_groups(const gpsiq_chan_desc* __restrict__ desc,
                    const CarrSpec* __restrict__ specE, const CarrSpec* __restrict__ specG, double* __restrict__ carr_ck,
                    size_t ck_plane, CarrInfo* __restrict__ infoG, size_t info_plane, const double* __restrict__ traceG,
                    double* __restrict__ carr_trace, GroupInfo* __restrict__ ginfo, int* __restrict__ fallbacks,
                    const TieEvent* __restrict__ tieG, const double* __restrict__ startS, const SliceRes* __restrict__ sres, const double* __restrict__ start,
                    const double* __restrict__ end, int* __restrict__ err, int E, int C, int N, int T, int ntiles) {
    __shared__ GroupEpoch s_ge[GROUP_EPOCHS];
    const int lane = threadIdx.x;
    const int ngroups = (E + GROUP_EPOCHS - 1) / GROUP_EPOCHS;
    const int g = blockIdx.x / C, c = blockIdx.x - g * C;
    if (g >= ngroups) return;
    const SliceRes res = sres[c];
    if (res.how == 0) return;                                       // chained serially by k_carr_final
    const double* sS = startS + (size_t) res.variant * ngroups * C + c;
    // the slice-level first wrap lies inside group 0 (slice_chain_group), so every later group start is translatable
    // (the groups after the one holding the trajectory's tie event start from the second translation: TieEvent)
    const double x_start = (g == 0 || res.how == 2) ? start[c] : add_rn(sS[(size_t) g * C], g > res.tie_g ? res.diff2 : res.diff);
    int fb = 0;
    const double x = final_group_warp(s_ge, x_start, g, desc, specE, specG, tieG, carr_ck, ck_plane, infoG, info_plane, traceG,
                                      carr_trace, ginfo, fb, E, C, c, N, T, ntiles, lane);
    if (lane == 0) {
        const double want = (g + 1 < ngroups) ? (res.how == 2 ? start[c] : add_rn(sS[(size_t) (g + 1) * C], g + 1 > res.tie_g ? res.diff2 : res.diff))
                                              : end[c];
        if (x != want) atomicOr(err, GPSIQ_DEVERR_SLICE);
        if (fb) atomicAdd(fallbacks, fb);
    }
}

// ---------------------------------------------------------------------------
// k_synth_lanes: warp = one tile of samples, lane = channel slot.
// Executes the very IEEE additions the reference executes, starting from the
// exact tile-start state; the across-satellite accumulate (plutogpssim.c:2705-2706)
// is a warp reduction.  Simple and exact by construction; kept as the
// cross-check for k_synth_line and as the path for epochs outside its contract.
// ---------------------------------------------------------------------------
#define LANES_WARPS 4
#define LANES_FLAGGED_CTAS 8   // CTAs per epoch of the fallback launch beside k_synth_line

__global__ void __launch_bounds__(LANES_WARPS * 32)
k_synth_lanes(const gpsiq_chan_desc* __restrict__ desc, const int2* __restrict__ lut,
              const double* __restrict__ code_ck, const int* __restrict__ wrap_ck,
              const CarrLookup carr, const uint32_t* __restrict__ ustart,
              const uint32_t* __restrict__ ca, const int* __restrict__ amp_sum, const int* __restrict__ step_flag,
              int only_flagged, int16_t* __restrict__ iq, int e0,
              int C, int N, int T, int ntiles, int tile_groups, int ctas_per_epoch, int carrier_mode) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int2* s_lut = reinterpret_cast<int2*>(smem_raw);                       // [C][512]
    uint32_t* s_ca = reinterpret_cast<uint32_t*>(s_lut + (size_t) C * 512); // [C][33]

    // ctas_per_epoch CTAs per epoch stride over its tile groups (the fallback launch beside k_synth_line keeps the grid
    // small: normally every CTA returns at once)
    const int e = e0 + blockIdx.x / ctas_per_epoch;
    if (only_flagged && !(amp_sum[e] > 32767 || step_flag[e])) return;  // rendered by k_synth_line
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const gpsiq_chan_desc* de = desc + (size_t) e * C;

    for (int i = threadIdx.x; i < C * 512; i += blockDim.x) s_lut[i] = lut[(size_t) e * C * 512 + i];
    for (int i = threadIdx.x; i < C * CA_WORDS; i += blockDim.x) {
        const int c = i / CA_WORDS, w = i - c * CA_WORDS;
        const int prn = de[c].prn;
        s_ca[c * 33 + w] = (prn > 0 && prn <= 32) ? ca[prn * CA_WORDS + w] : 0u;
    }
    __syncthreads();

  for (int tg = blockIdx.x % ctas_per_epoch; tg < tile_groups; tg += ctas_per_epoch) {
    const int t = tg * LANES_WARPS + warp;
    if (t >= ntiles) continue;
    const int n0 = t * T;
    const int len = min(T, N - n0);

    const bool active = (lane < C) && (de[lane < C ? lane : 0].prn > 0);
    double cp = 0.0, ph = 0.0, cstep = 0.0, pstep = 0.0;
    uint32_t uph = 0, ustep = 0;
    uint64_t navbits = 0;
    int icode = 0, kbit = 0;
    if (active) {
        const gpsiq_chan_desc d = de[lane];
        const size_t o = ((size_t) e * ntiles + t) * C + lane;
        cp = code_ck[o];
        const int w = wrap_ck[o] + d.ms0 % 20;
        kbit = w / 20;
        icode = w - kbit * 20;
        cstep = d.code_step;
        pstep = d.carr_step;
        ustep = (uint32_t) (int32_t) d.carr_step;
        if (carrier_mode == GPSIQ_CARRIER_FLOAT) ph = carr_lookup(carr, e, lane, t, T, N, C, ntiles);
        else uph = ustart[(size_t) e * C + lane] + ustep * (uint32_t) n0;   // closed form (plutogpssim.c:2748)
        navbits = d.navbits;
    }
    const int2* my_lut = s_lut + (size_t) (active ? lane : 0) * 512;
    const uint32_t* my_ca = s_ca + (active ? lane : 0) * 33;
    uint32_t* out = reinterpret_cast<uint32_t*>(iq) + (size_t) e * N + n0;

    uint32_t keep = 0;
    for (int n = 0; n < len; n++) {
        int vi = 0, vq = 0;
        if (active) {
            int it;
            if (carrier_mode == GPSIQ_CARRIER_FLOAT)
                it = min(__double2int_rd(__dmul_rn(ph, 512.0)), 511);  // plutogpssim.c:2697 (511 clamp: ph==1.0 corner)
            else
                it = (int) ((uph >> 16) & 0x1ff);                       // plutogpssim.c:2699
            const int chipi = __double2int_rz(cp);                      // plutogpssim.c:2737
            const uint32_t chip = (my_ca[chipi >> 5] >> (chipi & 31)) & 1u;
            const uint32_t nav = (uint32_t) (navbits >> kbit) & 1u;     // plutogpssim.c:2732
            const int2 a = my_lut[it];
            const bool neg = (chip != nav);                             // (2b-1)(2c-1) = +1 iff b == c
            vi = neg ? -a.x : a.x;
            vq = neg ? -a.y : a.y;
            // code NCO + NAV counters, plutogpssim.c:2709-2734
            cp = __dadd_rn(cp, cstep);
            if (cp >= 1023.0) {
                cp = __dadd_rn(cp, -1023.0);
                if (++icode >= 20) { icode = 0; kbit++; }
            }
            // carrier NCO, plutogpssim.c:2741-2748
            if (carrier_mode == GPSIQ_CARRIER_FLOAT) {
                ph = __dadd_rn(ph, pstep);
                if (ph >= 1.0) ph = __dadd_rn(ph, -1.0);
                else if (ph < 0.0) ph = __dadd_rn(ph, 1.0);
            } else {
                uph += ustep;
            }
        }
        const int si = __reduce_add_sync(0xffffffffu, vi);
        const int sq = __reduce_add_sync(0xffffffffu, vq);
        // (short) casts of plutogpssim.c:2754-2755: keep the low 16 bits of each sum
        const uint32_t word = ((uint32_t) si & 0xffffu) | ((uint32_t) sq << 16);
        if ((n & 31) == lane) keep = word;
        if ((n & 31) == 31 || n == len - 1) {
            const int base = n & ~31;
            if (base + lane <= n) out[base + lane] = keep;  // 128 B per warp, coalesced
        }
    }
  }
}

// ---------------------------------------------------------------------------
// k_checksum: order-independent per-epoch checksum of an I/Q stream
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    return x;
}

__global__ void k_checksum(const uint32_t* __restrict__ iq, unsigned long long* __restrict__ sums, int N) {
    const int e = blockIdx.y;
    const uint32_t* p = iq + (size_t) e * N;
    unsigned long long acc = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
        acc += mix64(((unsigned long long) (uint32_t) i << 32) | p[i]);
    for (int o = 16; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sums[e], acc);
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static void use_set(gpsiq_ctx* ctx, int i) {
    const ScanSet& ss = ctx->sets[i];
    ctx->d_lut = ss.d_lut; ctx->d_lutp = ss.d_lutp; ctx->d_flags = ss.d_flags; ctx->d_code_ck = ss.d_code_ck;
    ctx->d_wrap_ck = ss.d_wrap_ck; ctx->d_carr_ck = ss.d_carr_ck; ctx->d_drift = ss.d_drift;
    ctx->d_spec = ss.d_spec; ctx->d_specE = ss.d_specE; ctx->d_cinfo = ss.d_cinfo; ctx->d_info = ss.d_info;
    ctx->d_specG = ss.d_specG; ctx->d_ginfo = ss.d_ginfo; ctx->d_traceG = ss.d_traceG;
    ctx->d_specS = ss.d_specS; ctx->d_startS = ss.d_startS; ctx->d_sres = ss.d_sres; ctx->d_start0 = ss.d_start0;
    ctx->d_tieG = ss.d_tieG; ctx->d_tieS = ss.d_tieS;
    ctx->d_adv = ss.d_adv; ctx->d_carr_trace = ss.d_carr_trace; ctx->d_ustart = ss.d_ustart;
    ctx->d_lrecs = ss.d_lrecs; ctx->d_elist = ss.d_elist; ctx->d_hazlist = ss.d_hazlist;
    ctx->d_line_counters = ss.d_line_counters; ctx->d_patches = ss.d_patches;
    ctx->set_cur = i;
}

extern "C" {

const char* gpsiq_version(void) { return "gpsiq 0.1 (sm_100a)"; }

const char* gpsiq_strerror(int s) {
    switch (s) {
        case GPSIQ_OK: return "ok";
        case GPSIQ_ERR_ARG: return "bad argument";
        case GPSIQ_ERR_CUDA: return "CUDA error";
        case GPSIQ_ERR_NOMEM: return "out of memory";
        case GPSIQ_ERR_CAPACITY: return "batch exceeds max_epochs";
        default: return "unknown status";
    }
}

const char* gpsiq_last_error(const gpsiq_ctx* ctx) { return ctx ? ctx->err : g_err; }

void gpsiq_get_tables(int32_t* s, int32_t* c) {
    for (int i = 0; i < 512; i++) { s[i] = k_sin512[i]; c[i] = k_cos512[i]; }
}

int gpsiq_get_ca_code(int prn, uint8_t* chips) {
    if (prn < 1 || prn > 32 || !chips) return GPSIQ_ERR_ARG;
    ca_generate(prn, chips);
    return GPSIQ_OK;
}

int gpsiq_make_desc(gpsiq_chan_desc* out, int carrier_mode, int prn, double f_carr, double f_code, double delt,
                    double carr_phase, double code_phase, const uint64_t* dwrd60, int iword, int ibit, int icode,
                    double gain, int carr_phase_is_new) {
    return gpsiq_make_desc_inline(out, carrier_mode, prn, f_carr, f_code, delt, carr_phase, code_phase, dwrd60, iword, ibit,
                                  icode, gain, carr_phase_is_new);
}

int gpsiq_nco_advance(int mode, double* phase, double step, int64_t count, int64_t* wraps) {
    if (!phase || count < 0 || (mode != NCO_CODE && mode != NCO_CARRIER)) return GPSIQ_ERR_ARG;
    double x = *phase;
    int w = 0;
    int64_t wtot = 0;
    const StepInfo tab = step_info(step);
    while (count > 0) {
        const int chunk = count > (1 << 30) ? (1 << 30) : (int) count;
        w = 0;
        if (mode == NCO_CODE) nco_advance<NCO_CODE>(x, step, tab, chunk, w);
        else nco_advance<NCO_CARRIER>(x, step, tab, chunk, w);
        wtot += w;
        count -= chunk;
    }
    *phase = x;
    if (wraps) *wraps += wtot;
    return GPSIQ_OK;
}

// use_slice: chain through level 5 as the device does (slice-level speculation from the estimated start, one exact head
// scan, the groups chained independently from their translated start phases).  how_out: 1 translated, 0 serial,
// -2 = the groups' ends disagreed with the translation (an internal error: must never happen).
// study (may be NULL): group size, a residual rate subtracted from every closed-form epoch advance (what the device's
// k_prepare does with its measured rate), and where to report the margins (gpsiq_carrier_study_host).
struct HostStudy { int group_epochs; double est_rate; double* out; const int* flags; const double* phase0; };   // flags[e]: bit 0 = slot inactive in epoch e, bit 1 = re-seeded with phase0[e]

static int carrier_chain_host_impl(const double* steps, int n_epochs, int N, int T, double x0, double est_err, double* ck_out,
                                   double* x_end_out, int* n_fallback, int use_slice, int* how_out, int* ties_out,
                                   const HostStudy* study = NULL) {
    if (!steps || !ck_out || n_epochs < 1 || N < 1 || T < 1) return GPSIQ_ERR_ARG;
    const int ntiles = (N + T - 1) / T;
    const int G = (ntiles + 7) / 8, J = (ntiles + G - 1) / G;
    const int GP = (study && study->group_epochs > 0) ? study->group_epochs : 4;   // epochs per group (the device: 64)
    const double est_rate = study ? study->est_rate : 0.0;
    const int E = n_epochs;
    const size_t ep = (size_t) 7 * ntiles;              // planes of one epoch: [7][ntiles]
    double* planes = (double*) malloc(sizeof(double) * ep * E);
    GroupEpoch* ge = (GroupEpoch*) malloc(sizeof(GroupEpoch) * E);
    ChunkInfo* ci = (ChunkInfo*) malloc(sizeof(ChunkInfo) * 2 * J * E);
    CarrInfo* infG = (CarrInfo*) malloc(sizeof(CarrInfo) * 3 * E);   // [0],[1] group variants, [2] exact
    double* trace = (double*) malloc(sizeof(double) * 3 * E);
    double* est = (double*) malloc(sizeof(double) * E);
    CarrSpec* cs_all = (CarrSpec*) malloc(sizeof(CarrSpec) * 2 * SPEC_MAX_CHUNKS * E);   // chunk results of every epoch
    if (!planes || !ge || !ci || !infG || !trace || !est || !cs_all) return GPSIQ_ERR_NOMEM;
    int fb = 0;
    double xe = x0;
    // levels 1 + 2 per epoch: chunk speculation, stitch
    for (int e = 0; e < E; e++) {
        const double d = steps[e];
        GroupEpoch& g = ge[e];
        g.d = d; g.phase0 = 0.0; g.active = 1; g.reset = 0;
        if (study && study->flags) {
            g.active = !(study->flags[e] & 1);
            g.reset = g.active && (study->flags[e] & 2) != 0;
            if (g.reset) { g.phase0 = study->phase0 ? study->phase0[e] : 0.0; xe = g.phase0; }   // (the device: ereset)
        }
        const StepInfo si = step_info(d);
        // the epoch's closed-form advance modulo one cycle, as k_prepare computes it (exact product split: N*d is ~300
        // cycles, its double alone carries only ~5e-14 of absolute precision)
        const double dest = carr_drift_estimate(d, si, N);
        const double eadv = fma((double) N, d, dest);    // whole-epoch advance (chunk interpolation)
        volatile double pbig = (double) N * d;
        const double pe = fma((double) N, d, -pbig);
        const double eadv_frac = (pbig - floor(pbig)) + ((pe + dest) - est_rate);
        double a0 = xe + (g.reset ? 0.0 : est_err);      // what the device would guess, plus injected error
        a0 -= floor(a0);
        est[e] = a0;
        double* pl = planes + ep * e;
        CarrSpec* cs = cs_all + (size_t) e * 2 * SPEC_MAX_CHUNKS;
        for (int j = 0; j < J; j++) {
            CarrSpec& o0 = cs[j * 2];
            CarrSpec& o1 = cs[j * 2 + 1];
            o0.margin = -1.0; o0.n1 = -1; o0.xw1 = 0; o0.xend = 0; o0.pad = 0;
            o1 = o0;
            if (!g.active || !carr_step_speculable(d)) continue;
            const int t0 = j * G, t1 = (t0 + G < ntiles) ? t0 + G : ntiles;
            double xs = a0;
            if (j > 0) { xs = a0 + eadv * ((double) (t0 * T) / (double) N); xs -= floor(xs); if (!(xs >= 0.0 && xs < 1.0)) xs = 0.0; }
            spec_scan_chunk(xs, d, si, N, T, t0, t1, pl, pl + (size_t) ntiles, 1, o0, o1);
        }
        CarrSpec* sE[2] = {&g.s0, &g.s1};
        for (int V = 0; V < 2; V++) {
            sE[V]->margin = -1.0; sE[V]->n1 = -1; sE[V]->xw1 = 0; sE[V]->xend = 0; sE[V]->pad = 0;
            if (!g.active || (V == 1 && d >= 0.0) || !carr_step_speculable(d)) continue;
            stitch_epoch(a0, d, si, N, T, G, V, cs, pl + (size_t) (2 + V) * ntiles, 1, ci + ((size_t) e * 2 + V) * J, *sE[V]);
        }
        if (!g.active) continue;                          // (the phase passes through an inactive epoch)
        double t2 = xe + eadv_frac;
        t2 -= floor(t2);
        xe = (t2 >= 0.0 && t2 < 1.0) ? t2 : 0.0;
    }
    // level 3 per group
    const int ngroups = (E + GP - 1) / GP;
    CarrSpec* sGall = (CarrSpec*) malloc(sizeof(CarrSpec) * 2 * ngroups);
    double* startS = (double*) malloc(sizeof(double) * 2 * ngroups);
    TieEvent* tGall = (TieEvent*) malloc(sizeof(TieEvent) * 2 * ngroups);
    if (!sGall || !startS || !tGall) return GPSIQ_ERR_NOMEM;
    int ties_seen = 0, ties_applied = 0;
    for (int first = 0; first < E; first += GP) {
        const int count = (E - first < GP) ? E - first : GP;
        CarrSpec* sG = sGall + 2 * (first / GP);
        bool any_neg = false;
        for (int k = 0; k < count; k++) any_neg |= ge[first + k].active && ge[first + k].d < 0.0;
        for (int V = 0; V < 2; V++) {
            sG[V].margin = -1.0; sG[V].n1 = -1; sG[V].xw1 = 0; sG[V].xend = 0; sG[V].pad = 0;
            TieEvent& tG = tGall[2 * (first / GP) + V];
            tG.pos = -1; tG.k = 0;
            if (V == 1 && !any_neg) continue;
            group_chain(est[first], ge + first, count, N, T, V, planes + ep * first + (size_t) (4 + V) * ntiles, 1, ep,
                        infG + (size_t) V * E + first, 1, trace + (size_t) V * E + first, 1, sG[V], tG, fb);
            if (tG.pos >= 0 && sG[V].margin > 0.0) ties_seen++;
        }
    }
    // level 5: slice-level speculation from the estimated start, then the exact head scan
    int how = 0, vS = 0, tie_g = 0x7fffffff;
    double diffS = 0.0, diffS2 = 0.0, x_end_slice = 0.0;
    if (use_slice) {
        CarrSpec sS[2];
        TieEvent tS[2];
        for (int V = 0; V < 2; V++) {
            GroupTrack tr;
            tr.margin = 1.0; tr.xw1 = 0.0; tr.pos = -1; tr.usable = 1; tr.tie.pos = -1; tr.tie.k = 0;
            double xs = est[0];
            for (int g = 0; g < ngroups && tr.usable; g++) {
                const int first = g * GP, count = (E - first < GP) ? E - first : GP;
                startS[V * ngroups + g] = xs;
                // as the device does it: the group's first epoch alone, then the rest
                int r = slice_chain_group(xs, ge + first, 1, N, sGall[2 * g], sGall[2 * g + 1], tGall[2 * g], tGall[2 * g + 1],
                                          tr, V, first, g, 0);
                if (r == 0 && count > 1)
                    r = slice_chain_group(xs, ge + first + 1, count - 1, N, sGall[2 * g], sGall[2 * g + 1], tGall[2 * g],
                                          tGall[2 * g + 1], tr, V, first, g, 1);
                if (r < 0) tr.usable = 0;
            }
            sS[V].xw1 = tr.xw1; sS[V].xend = xs; sS[V].n1 = tr.pos; sS[V].pad = 0;
            sS[V].margin = (tr.pos >= 0 && tr.usable) ? tr.margin : -1.0;
            tS[V] = tr.tie;
            if (study && study->out) study->out[V] = sS[V].margin;
        }
        double xv = x0;
        const int count0 = (E < GP) ? E : GP;
        int r = slice_verify(xv, ge, 1, N, sS[0], sS[1], tS[0], tS[1], vS, diffS, diffS2, 0);
        if (r == 0 && count0 > 1) r = slice_verify(xv, ge + 1, count0 - 1, N, sS[0], sS[1], tS[0], tS[1], vS, diffS, diffS2, 1);
        if (r == 1) {
            how = 1; x_end_slice = xv;
            if (tS[vS].pos >= 0) tie_g = tS[vS].pos;
            if (diffS2 != diffS) ties_applied += 1000;
        }
    }
    // level 4 per group: serially, or -- after a slice-level match -- every group from its own translated start phase
    double x = x0;
    for (int first = 0; first < E; first += GP) {
        const int count = (E - first < GP) ? E - first : GP;
        const CarrSpec* sG = sGall + 2 * (first / GP);
        const TieEvent* tG = tGall + 2 * (first / GP);
        if (how == 1 && first > 0) {
            const double xs = add_rn(startS[vS * ngroups + first / GP], first / GP > tie_g ? diffS2 : diffS);
            if (xs != x) how = -2;     // the previous group did not end where the translation says this one starts
            x = xs;
        }
        GroupInfo gi;
        x = group_final(x, ge + first, count, N, T, sG[0], sG[1], planes + ep * first + (size_t) 6 * ntiles, 1, ep,
                        infG + (size_t) 2 * E + first, 1, trace + first, trace + (size_t) E + first,
                        trace + (size_t) 2 * E + first, 1, gi, fb, tG[0], tG[1]);
        if (gi.pos != 0x7fffffff && gi.delta2 != gi.delta) ties_applied++;
        for (int k = 0; k < count; k++) {
            const int e = first + k;
            for (int t = 0; t < ntiles; t++)
                ck_out[(size_t) e * ntiles + t] = carr_tile_phase(planes + ep * e, (size_t) ntiles, 1, t, T, N, G, J, k, gi,
                                                                  infG[e], infG[(size_t) E + e], infG[(size_t) 2 * E + e],
                                                                  ci + (size_t) e * 2 * J, cs_all + (size_t) e * 2 * SPEC_MAX_CHUNKS);
        }
    }
    if (how == 1 && x != x_end_slice) how = -2;
    if (study && study->out) {
        double e = xe - x;                               // closed-form end estimate (from the exact start) - exact end
        e -= rint(e);
        study->out[2] = e;
        double mg = 1.0;                                 // smallest usable group-level margin (variant 0)
        int unusable = 0;
        for (int g = 0; g < ngroups; g++) { const double m = sGall[2 * g].margin; if (m > 0.0) { if (m < mg) mg = m; } else unusable++; }
        study->out[3] = mg;
        study->out[4] = (double) unusable;
    }
    free(planes); free(ge); free(ci); free(infG); free(trace); free(est); free(cs_all); free(sGall); free(startS); free(tGall);
    if (x_end_out) *x_end_out = x;
    if (n_fallback) *n_fallback = fb;
    if (how_out) *how_out = how;
    if (ties_out) { ties_out[0] = ties_seen; ties_out[1] = ties_applied; }
    return GPSIQ_OK;
}

int gpsiq_carrier_chain_host(const double* steps, int n_epochs, int N, int T, double x0, double est_err, double* ck_out,
                             double* x_end_out, int* n_fallback) {
    return carrier_chain_host_impl(steps, n_epochs, N, T, x0, est_err, ck_out, x_end_out, n_fallback, 0, NULL, NULL);
}

int gpsiq_carrier_study_host(const double* steps, int n_epochs, int N, int T, double x0, double est_err, double est_rate,
                             int group_epochs, double* out5, int* how_out, int* n_fallback) {
    if (!out5) return GPSIQ_ERR_ARG;
    const int ntiles = (N + T - 1) / T;
    double* ck = (double*) malloc(sizeof(double) * (size_t) n_epochs * ntiles);
    if (!ck) return GPSIQ_ERR_NOMEM;
    for (int i = 0; i < 5; i++) out5[i] = 0.0;
    HostStudy st;
    st.group_epochs = group_epochs; st.est_rate = est_rate; st.out = out5; st.flags = NULL; st.phase0 = NULL;
    const int rc = carrier_chain_host_impl(steps, n_epochs, N, T, x0, est_err, ck, NULL, n_fallback, 1, how_out, NULL, &st);
    free(ck);
    return rc;
}

int gpsiq_carrier_slice_host(const double* steps, int n_epochs, int N, int T, double x0, double est_err, double* ck_out,
                             double* x_end_out, int* n_fallback, int* how_out, int* ties_out, const int* flags,
                             const double* phase0) {
    HostStudy st;
    st.group_epochs = 0; st.est_rate = 0.0; st.out = NULL; st.flags = flags; st.phase0 = phase0;
    return carrier_chain_host_impl(steps, n_epochs, N, T, x0, est_err, ck_out, x_end_out, n_fallback, 1, how_out, ties_out,
                                   flags ? &st : NULL);
}

void* gpsiq_host_alloc(size_t bytes) {
    void* p = NULL;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return NULL;
    return p;
}
void gpsiq_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

// Everything of gpsiq_create that can fail after the context exists: on failure the caller copies the message to
// the global error text (gpsiq_last_error(NULL)) and destroys the partly built context.
static int create_body(gpsiq_ctx* ctx, const gpsiq_config* cfg, const cudaDeviceProp& prop) {
    ctx->cfg = *cfg;
    ctx->trace_on = getenv("GPSIQ_TRACE") != NULL;
    if (getenv("GPSIQ_LINE_GRID_CAP")) ctx->line_grid_cap = atoi(getenv("GPSIQ_LINE_GRID_CAP"));  // experiments
    ctx->render_waits_spec = getenv("GPSIQ_RENDER_WAITS_SPEC") ? atoi(getenv("GPSIQ_RENDER_WAITS_SPEC")) : 0;
    if (ctx->trace_on) ctx->trace = (TraceRec*) calloc(TRACE_MAX, sizeof(TraceRec));
    ctx->sm_count = prop.multiProcessorCount;
    ctx->C = cfg->max_chan;
    ctx->N = cfg->samples_per_epoch;
    ctx->E = cfg->max_epochs;
    ctx->T = cfg->tile_samples ? cfg->tile_samples : LN_TILE;
    ctx->T = (ctx->T + 31) & ~31;
    ctx->use_line = (cfg->kernel == GPSIQ_KERNEL_AUTO || cfg->kernel == GPSIQ_KERNEL_LINE) && ctx->T == LN_TILE &&
                    ctx->C <= 32;
    if (cfg->kernel == GPSIQ_KERNEL_LINE && !ctx->use_line)
        return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_create: the line kernel needs tile_samples 0/1024 and max_chan <= 32",
                    cudaSuccess);
    ctx->ntiles = (ctx->N + ctx->T - 1) / ctx->T;
    CU(cudaSetDevice(cfg->device));
    CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&ctx->ev_sub[0], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ctx->ev_sub[1], cudaEventDisableTiming));
    {   // the patch walk of the batch being rendered runs beside its sample kernel and must be done when that ends:
        // highest priority, so that its few blocks get SM slots at once
        int plo = 0, phi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&plo, &phi));
        CU(cudaStreamCreateWithPriority(&ctx->aux_stream, cudaStreamNonBlocking, phi));
    }
    CU(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ctx->ev_chain, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ctx->ev_fork2, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ctx->ev_code2, cudaEventDisableTiming));
    for (int i = 0; i < 2; i++) {
        CU(cudaEventCreateWithFlags(&ctx->ev_P[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_F[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < TIMING_RING; i++)
        for (int j = 0; j < 5; j++) CU(cudaEventCreate(&ctx->ev[i][j]));
    const size_t EC = (size_t) ctx->E * ctx->C;
    const size_t ck = EC * ctx->ntiles;
    CU(cudaMalloc(&ctx->d_desc, EC * sizeof(gpsiq_chan_desc)));
    ctx->ck_plane = ck;
    {   // chunks per epoch of the first speculation level (default 8; GPSIQ_SPEC_CHUNKS: experiments)
        int chunks = 8;
        if (getenv("GPSIQ_SPEC_CHUNKS")) chunks = atoi(getenv("GPSIQ_SPEC_CHUNKS"));
        if (chunks < 1) chunks = 1;
        if (chunks > SPEC_MAX_CHUNKS) chunks = SPEC_MAX_CHUNKS;
        ctx->G = (ctx->ntiles + chunks - 1) / chunks;
    }
    ctx->J = (ctx->ntiles + ctx->G - 1) / ctx->G;
    ctx->slice_spec = !(getenv("GPSIQ_SLICE_SPEC") && atoi(getenv("GPSIQ_SLICE_SPEC")) == 0);
    for (int i = 0; i < NSETS; i++) {
        ScanSet& ss = ctx->sets[i];
        CU(cudaMalloc(&ss.d_descbuf, EC * sizeof(gpsiq_chan_desc)));
        CU(cudaMalloc(&ss.d_lut, EC * 512 * sizeof(int2)));
        CU(cudaMalloc(&ss.d_lutp, EC * 512 * sizeof(int32_t)));
        CU(cudaMalloc(&ss.d_flags, 2 * (size_t) ctx->E * sizeof(int)));
        CU(cudaMemset(ss.d_flags, 0, 2 * (size_t) ctx->E * sizeof(int)));
        CU(cudaMalloc(&ss.d_code_ck, ck * sizeof(double)));
        CU(cudaMalloc(&ss.d_wrap_ck, ck * sizeof(int)));
        CU(cudaMalloc(&ss.d_carr_ck, 7 * ck * sizeof(double)));
        CU(cudaMalloc(&ss.d_specE, EC * 2 * sizeof(CarrSpec)));
        CU(cudaMalloc(&ss.d_cinfo, EC * 2 * ctx->J * sizeof(ChunkInfo)));
        CU(cudaMalloc(&ss.d_drift, 4 * EC * sizeof(double)));
        CU(cudaMalloc(&ss.d_spec, EC * 2 * ctx->J * sizeof(CarrSpec)));
        CU(cudaMalloc(&ss.d_info, 3 * EC * sizeof(CarrInfo)));
        {
            const size_t ng = ((size_t) ctx->E + GROUP_EPOCHS - 1) / GROUP_EPOCHS;
            CU(cudaMalloc(&ss.d_specG, ng * ctx->C * 2 * sizeof(CarrSpec)));
            CU(cudaMalloc(&ss.d_ginfo, ng * ctx->C * sizeof(GroupInfo)));
            CU(cudaMalloc(&ss.d_traceG, 2 * EC * sizeof(double)));
            CU(cudaMalloc(&ss.d_specS, (size_t) ctx->C * 2 * sizeof(CarrSpec)));
            CU(cudaMalloc(&ss.d_startS, 2 * ng * ctx->C * sizeof(double)));
            CU(cudaMalloc(&ss.d_sres, (size_t) ctx->C * sizeof(SliceRes)));
            CU(cudaMemset(ss.d_sres, 0, (size_t) ctx->C * sizeof(SliceRes)));
            CU(cudaMalloc(&ss.d_start0, ctx->C * sizeof(double)));
            CU(cudaMalloc(&ss.d_rate_used, ctx->C * sizeof(double)));
            CU(cudaMalloc(&ss.d_ccum_used, ctx->C * sizeof(double)));
            CU(cudaMemset(ss.d_ccum_used, 0, ctx->C * sizeof(double)));
            CU(cudaMemset(ss.d_rate_used, 0, ctx->C * sizeof(double)));
            CU(cudaMalloc(&ss.d_tieG, ng * ctx->C * 2 * sizeof(TieEvent)));
            CU(cudaMalloc(&ss.d_tieS, (size_t) ctx->C * 2 * sizeof(TieEvent)));
        }
        CU(cudaMalloc(&ss.d_adv, 2 * ctx->C * sizeof(double)));
        CU(cudaMalloc(&ss.d_carr_trace, EC * sizeof(double)));
        CU(cudaMalloc(&ss.d_ustart, EC * sizeof(uint32_t)));
        CU(cudaMalloc(&ss.d_est, ctx->C * sizeof(double)));
        CU(cudaMalloc(&ss.d_exact_end, ctx->C * sizeof(double)));
        CU(cudaEventCreateWithFlags(&ss.adv_done, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ss.est_done, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ss.spec_all, cudaEventDisableTiming));
        ss.seq = -1;
        CU(cudaEventCreateWithFlags(&ss.scan_done, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ss.render_done, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ss.spec_done, cudaEventDisableTiming));
        ss.n_epochs = 0;
    }
    use_set(ctx, 0);
    {
        // the scan streams carry small latency-bound kernels that must get onto the SMs beside the sample
        // kernel of the previous batch: highest priority
        int prio_lo = 0, prio_hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        const int scan_prio = (getenv("GPSIQ_SCAN_PRIO") && !strcmp(getenv("GPSIQ_SCAN_PRIO"), "lo")) ? prio_lo : prio_hi;  // experiments
        CU(cudaStreamCreateWithPriority(&ctx->scan_stream, cudaStreamNonBlocking, scan_prio));
        CU(cudaStreamCreateWithPriority(&ctx->aux2_stream, cudaStreamNonBlocking, prio_hi));
        for (int i = 0; i < NSETS; i++) CU(cudaStreamCreateWithPriority(&ctx->sets[i].stream, cudaStreamNonBlocking, scan_prio));
        CU(cudaEventCreateWithFlags(&ctx->ev_final, cudaEventDisableTiming));
    }
    CU(cudaMalloc(&ctx->d_fallbacks, sizeof(int)));
    CU(cudaMemset(ctx->d_fallbacks, 0, sizeof(int)));
    CU(cudaMalloc(&ctx->d_est_meas, ctx->C * sizeof(double)));
    CU(cudaMemset(ctx->d_est_meas, 0, ctx->C * sizeof(double)));
    CU(cudaMalloc(&ctx->d_est_ccum, ctx->C * sizeof(double)));
    CU(cudaMemset(ctx->d_est_ccum, 0, ctx->C * sizeof(double)));
    CU(cudaMalloc(&ctx->d_slice_stats, 2 * sizeof(unsigned long long)));
    CU(cudaMemset(ctx->d_slice_stats, 0, 2 * sizeof(unsigned long long)));
    CU(cudaMalloc(&ctx->d_carr_state, ctx->C * sizeof(double)));
    CU(cudaMalloc(&ctx->d_est_state, ctx->C * sizeof(double)));
    CU(cudaMemset(ctx->d_est_state, 0, ctx->C * sizeof(double)));
    CU(cudaMalloc(&ctx->d_bias_rate, ctx->C * sizeof(double)));
    CU(cudaMemset(ctx->d_bias_rate, 0, ctx->C * sizeof(double)));
    CU(cudaMalloc(&ctx->d_carr_start, ctx->C * sizeof(double)));
    CU(cudaMalloc(&ctx->d_adv_prev, 2 * ctx->C * sizeof(double)));
    CU(cudaMalloc(&ctx->d_ca, 33 * CA_WORDS * sizeof(uint32_t)));
    CU(cudaMalloc(&ctx->d_iq, (size_t) ctx->E * ctx->N * 4));
    CU(cudaMalloc(&ctx->d_sums, (size_t) ctx->E * sizeof(unsigned long long)));
    CU(cudaMalloc(&ctx->d_err, sizeof(int)));
    CU(cudaMemset(ctx->d_carr_state, 0, ctx->C * sizeof(double)));
    CU(cudaMemset(ctx->d_err, 0, sizeof(int)));
    CU(cudaMemcpyToSymbol(c_sin512, k_sin512, sizeof k_sin512));
    CU(cudaMemcpyToSymbol(c_cos512, k_cos512, sizeof k_cos512));
    uint32_t h_ca[33 * CA_WORDS];
    memset(h_ca, 0, sizeof h_ca);
    for (int prn = 1; prn <= 32; prn++) {
        uint8_t chips[GPSIQ_CA_LEN];
        ca_generate(prn, chips);
        for (int i = 0; i < GPSIQ_CA_LEN; i++) h_ca[prn * CA_WORDS + (i >> 5)] |= (uint32_t) chips[i] << (i & 31);
    }
    CU(cudaMemcpy(ctx->d_ca, h_ca, sizeof h_ca, cudaMemcpyHostToDevice));
    if (ctx->use_line) {
        // The kernels meant to run beside k_synth_line (the rest of the next batch's carrier chain) ask for the
        // same (maximum) shared-memory carve-out: an SM cannot change its L1/shared split while blocks are
        // resident, so a kernel preferring another split would wait for the sample kernel's persistent CTAs
        // to leave.  No scan kernel keeps per-thread tables any more (nco_scan.cuh: binade_delta), so EVERY kernel
        // of the pipeline asks for this split and any of them can be placed beside the sample kernel.
#define CARVE(k) CU(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, (int) cudaSharedmemCarveoutMaxShared))
        CARVE(k_synth_line); CARVE(k_carr_stitch); CARVE(k_carr_group); CARVE(k_carr_final); CARVE(k_carr_slice); CARVE(k_carr_final_groups); CARVE(k_est_open_loop); CARVE(k_est_from_slice); CARVE(k_line_apply);
        CARVE(k_synth_lanes);
        CARVE(k_carr_speculate); CARVE(k_scan_code); CARVE(k_prepare); CARVE(k_line_anchor); CARVE(k_line_patch);
        CARVE(k_epoch_estimates); CARVE(k_slice_advance); CARVE(k_est_fold); CARVE(k_est_correct); CARVE(k_int_carrier);
        CARVE(k_bias_update); CARVE(k_int_fold); CARVE(k_checksum); CARVE(k_line_check); CARVE(k_line_refine);
#undef CARVE
        CU(cudaFuncSetAttribute(k_synth_line, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ln_smem_bytes(ctx->C)));
        const size_t tiles = (size_t) ctx->E * ctx->ntiles;
        for (int i = 0; i < NSETS; i++) CU(cudaMalloc(&ctx->d_anch[i], tiles * ctx->C * sizeof(ulonglong2)));
        const int dbg = cfg->reserved[1];
        ctx->haz_cap = (dbg & LN_DBG_FORCE_TILE) ? (int) (tiles * ctx->C) : (int) (tiles * ctx->C / 64 + 1024);
        ctx->patch_cap = 1 << 20;
        for (int i = 0; i < NSETS; i++) {   // per scan set: a batch's anchors / check / patch list are made right after its chain
            ScanSet& ss = ctx->sets[i];
            CU(cudaMalloc(&ss.d_hazlist, (size_t) ctx->haz_cap * 4));
            CU(cudaMalloc(&ss.d_lrecs, (size_t) ctx->E * ctx->C * sizeof(LineEpoch)));
            CU(cudaMalloc(&ss.d_elist, (size_t) ctx->E * ctx->C * 4));
            CU(cudaMalloc(&ss.d_patches, (size_t) ctx->patch_cap * sizeof(LinePatch)));
            CU(cudaMalloc(&ss.d_line_counters, 4 * sizeof(int)));
            CU(cudaEventCreateWithFlags(&ss.anchor_done, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&ss.anch_ready, cudaEventDisableTiming));
        }
        use_set(ctx, 0);
        CU(cudaMalloc(&ctx->d_line_totals, 4 * sizeof(unsigned long long)));
        CU(cudaMemset(ctx->d_line_totals, 0, 4 * sizeof(unsigned long long)));
        // chip/NAV sign tables: variant v = pol0*2 + pol1; entry k < 1023: chip k under NAV bit pol0,
        // k >= 1023: chip k-1023 of the NEXT code period under NAV bit pol1.  +1 iff NAV bit == chip
        // (plutogpssim.c:2701, 2732, 2737)
        int8_t* h4 = (int8_t*) malloc((size_t) 33 * 4 * LN_VS);
        if (!h4) return fail(ctx, GPSIQ_ERR_NOMEM, "gpsiq_create: out of host memory", cudaSuccess);
        memset(h4, 1, (size_t) 33 * 4 * LN_VS);
        for (int prn = 1; prn <= 32; prn++) {
            uint8_t chips[GPSIQ_CA_LEN];
            ca_generate(prn, chips);
            for (int v = 0; v < 4; v++)
                for (int k = 0; k < LN_VS; k++) {
                    const int pol = (k < GPSIQ_CA_LEN) ? (v >> 1) : (v & 1);
                    h4[((size_t) prn * 4 + v) * LN_VS + k] = (chips[k % GPSIQ_CA_LEN] == pol) ? 1 : -1;
                }
        }
        CU(cudaMalloc(&ctx->d_chips4, (size_t) 33 * 4 * LN_VS));
        cudaError_t ce3 = cudaMemcpy(ctx->d_chips4, h4, (size_t) 33 * 4 * LN_VS, cudaMemcpyHostToDevice);
        free(h4);
        CU(ce3);
    }
    const size_t smem_lanes = (size_t) ctx->C * 512 * sizeof(int2) + (size_t) ctx->C * 33 * 4;
    CU(cudaFuncSetAttribute(k_synth_lanes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_lanes));
    return GPSIQ_OK;
}

void gpsiq_destroy(gpsiq_ctx* ctx);

int gpsiq_create(gpsiq_ctx** out, const gpsiq_config* cfg) {
    gpsiq_ctx* ctx = NULL;
    if (!out || !cfg) return fail(NULL, GPSIQ_ERR_ARG, "gpsiq_create: null argument", cudaSuccess);
    *out = NULL;
    if (cfg->max_chan < 1 || cfg->max_chan > GPSIQ_MAX_CHAN || cfg->samples_per_epoch < 1 || cfg->max_epochs < 1 ||
        (cfg->carrier_mode != GPSIQ_CARRIER_FLOAT && cfg->carrier_mode != GPSIQ_CARRIER_INT32) || cfg->tile_samples < 0 ||
        (cfg->kernel != GPSIQ_KERNEL_AUTO && cfg->kernel != GPSIQ_KERNEL_LANE_PER_CHANNEL && cfg->kernel != GPSIQ_KERNEL_LINE))
        return fail(NULL, GPSIQ_ERR_ARG, "gpsiq_create: bad config", cudaSuccess);
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev < 1)
        return fail(NULL, GPSIQ_ERR_CUDA, "gpsiq_create: no CUDA device (there is no CPU fallback)", ce);
    if (cfg->device < 0 || cfg->device >= ndev) return fail(NULL, GPSIQ_ERR_ARG, "gpsiq_create: bad device", cudaSuccess);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) {
        snprintf(g_err, sizeof g_err, "gpsiq_create: device %d is sm_%d%d; this build is sm_100a only", cfg->device,
                 prop.major, prop.minor);
        return GPSIQ_ERR_CUDA;
    }
    ctx = (gpsiq_ctx*) calloc(1, sizeof *ctx);
    if (!ctx) return fail(NULL, GPSIQ_ERR_NOMEM, "gpsiq_create: out of host memory", cudaSuccess);
    const int rc = create_body(ctx, cfg, prop);
    if (rc != GPSIQ_OK) {
        snprintf(g_err, sizeof g_err, "%.255s", ctx->err);
        gpsiq_destroy(ctx);
        cudaGetLastError();
        return rc;
    }
    *out = ctx;
    return GPSIQ_OK;
}

int gpsiq_trace_dump(gpsiq_ctx* ctx, int reset);

void gpsiq_destroy(gpsiq_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->cfg.device);
    if (ctx->trace_on && getenv("GPSIQ_TRACE")[0] == '2') {  // GPSIQ_TRACE=2: dump the last records at destruction
        const int keep = 360;
        if (ctx->trace_n > keep) {  // keep the tail
            for (int i = 0; i < keep; i++) { TraceRec t = ctx->trace[i]; ctx->trace[i] = ctx->trace[ctx->trace_n - keep + i]; ctx->trace[ctx->trace_n - keep + i] = t; }
            ctx->trace_n = keep;
        }
        gpsiq_trace_dump(ctx, 1);
    }
    // (also called on a partly built context by gpsiq_create: every handle may still be NULL)
    cudaDeviceSynchronize();
    cudaFree(ctx->d_desc);
    for (int i = 0; i < NSETS; i++) {
        ScanSet& ss = ctx->sets[i];
        cudaFree(ss.d_descbuf); cudaFree(ss.d_lut); cudaFree(ss.d_lutp); cudaFree(ss.d_flags); cudaFree(ss.d_code_ck);
        cudaFree(ss.d_wrap_ck); cudaFree(ss.d_carr_ck); cudaFree(ss.d_drift); cudaFree(ss.d_spec);
        cudaFree(ss.d_specE); cudaFree(ss.d_cinfo); cudaFree(ss.d_info); cudaFree(ss.d_adv); cudaFree(ss.d_carr_trace);
        cudaFree(ss.d_specG); cudaFree(ss.d_ginfo); cudaFree(ss.d_traceG); cudaFree(ss.d_ustart);
        cudaFree(ss.d_specS); cudaFree(ss.d_startS); cudaFree(ss.d_sres); cudaFree(ss.d_start0);
        cudaFree(ss.d_tieG); cudaFree(ss.d_tieS); cudaFree(ss.d_rate_used); cudaFree(ss.d_ccum_used);
        if (ss.scan_done) cudaEventDestroy(ss.scan_done);
        if (ss.render_done) cudaEventDestroy(ss.render_done);
        if (ss.spec_done) cudaEventDestroy(ss.spec_done);
        if (ss.adv_done) cudaEventDestroy(ss.adv_done);
        if (ss.est_done) cudaEventDestroy(ss.est_done);
        if (ss.spec_all) cudaEventDestroy(ss.spec_all);
        if (ss.stream) cudaStreamDestroy(ss.stream);
        cudaFree(ss.d_est); cudaFree(ss.d_exact_end);
        cudaFree(ss.d_hazlist); cudaFree(ss.d_lrecs); cudaFree(ss.d_elist); cudaFree(ss.d_patches); cudaFree(ss.d_line_counters);
        if (ss.anchor_done) cudaEventDestroy(ss.anchor_done);
        if (ss.anch_ready) cudaEventDestroy(ss.anch_ready);
    }
    cudaFree(ctx->d_chips4);
    for (int i = 0; i < NSETS; i++) { cudaFree(ctx->d_anch[i]); if (ctx->h_stage[i]) cudaFreeHost(ctx->h_stage[i]); }
    cudaFree(ctx->d_line_totals);
    cudaFree(ctx->d_bias_rate); cudaFree(ctx->d_carr_start); cudaFree(ctx->d_adv_prev);
    if (ctx->d_mbox_peer) cudaIpcCloseMemHandle(ctx->d_mbox_peer);
    cudaFree(ctx->d_mbox);
    cudaFree(ctx->d_fallbacks); cudaFree(ctx->d_carr_state); cudaFree(ctx->d_est_state); cudaFree(ctx->d_ca);
    cudaFree(ctx->d_iq); cudaFree(ctx->d_iq2); cudaFree(ctx->d_sums); cudaFree(ctx->d_err); cudaFree(ctx->d_slice_stats); cudaFree(ctx->d_est_meas); cudaFree(ctx->d_est_ccum);
#define DROP_STREAM(s) do { if (s) cudaStreamDestroy(s); } while (0)
#define DROP_EVENT(e) do { if (e) cudaEventDestroy(e); } while (0)
    DROP_STREAM(ctx->scan_stream); DROP_STREAM(ctx->aux2_stream); DROP_EVENT(ctx->ev_fork2); DROP_EVENT(ctx->ev_code2);
    for (int i = 0; i < TIMING_RING; i++)
        for (int j = 0; j < 5; j++) DROP_EVENT(ctx->ev[i][j]);
    DROP_STREAM(ctx->stream); DROP_STREAM(ctx->copy_stream); DROP_STREAM(ctx->aux_stream);
    DROP_EVENT(ctx->ev_fork); DROP_EVENT(ctx->ev_chain); DROP_EVENT(ctx->ev_final);
    for (int i = 0; i < 2; i++) { DROP_EVENT(ctx->ev_P[i]); DROP_EVENT(ctx->ev_F[i]); }
    DROP_EVENT(ctx->ev_sub[0]); DROP_EVENT(ctx->ev_sub[1]);
    if (ctx->trace) {
        for (int i = 0; i < TRACE_MAX; i++) DROP_EVENT(ctx->trace[i].ev);
        free(ctx->trace);
    }
#undef DROP_STREAM
#undef DROP_EVENT
    cudaGetLastError();
    free(ctx);
}

// Phase 1a: amplitude LUTs, binade tables, contract flags, and the batch's closed-form phase advance.
// Every batch goes through a ring of two scan sets: begin_batch claims the free one (waiting, on
// the given stream, until its previous batch has been rendered), the scan phases fill it, the
// render phase consumes the oldest chained one.  A plain gpsiq_synth_device uses the ring with one
// batch in flight; submit/fetch and the time-slice runner keep two.
static int begin_batch(gpsiq_ctx* ctx, cudaStream_t st) {
    if (ctx->sets[ctx->set_prep].phase != 0)
        return fail(ctx, GPSIQ_ERR_CAPACITY, "every scan set is in flight (render a batch first)", cudaSuccess);
    CU(cudaStreamWaitEvent(st, ctx->sets[ctx->set_prep].render_done, 0));
    use_set(ctx, ctx->set_prep);
    return GPSIQ_OK;
}

static int enqueue_prepare(gpsiq_ctx* ctx, const gpsiq_chan_desc* desc_dev, int n_epochs, cudaStream_t st,
                           bool begun = false) {
    const int C = ctx->C, N = ctx->N;
    const int EC = n_epochs * C;
    if (!begun) { int rc0 = begin_batch(ctx, st); if (rc0) return rc0; }
    ScanSet& set = ctx->sets[ctx->set_prep];
    ctx->set_prep = (ctx->set_prep + 1) % NSETS;
    set.desc = desc_dev;
    set.n_epochs = n_epochs;
    set.phase = 1;
    set.anchored = 0;
    set.seq = ctx->seq++;
    if (ctx->ev_count < TIMING_RING && st != ctx->scan_stream) CU(cudaEventRecord(ctx->ev[ctx->ev_count][0], st));
    CU(cudaMemsetAsync(ctx->d_flags, 0, 2 * (size_t) ctx->E * sizeof(int), st));
    // the batch is prepared with a snapshot of the residual-rate estimate (k_bias_update needs to know which)
    CU(cudaMemcpyAsync(set.d_rate_used, ctx->d_bias_rate, C * sizeof(double), cudaMemcpyDeviceToDevice, st));
    k_prepare<<<EC, 128, 0, st>>>(desc_dev, ctx->d_lut, ctx->d_lutp, ctx->d_drift, ctx->d_flags,
                                  ctx->d_flags + ctx->E, C, N, ctx->cfg.carrier_mode, set.d_rate_used,
                                  ctx->d_drift + 3 * (size_t) ctx->E * C, ctx->d_err);
    ctx->launches += 1;
    trace_mark(ctx, st, "k_prepare");
    if (ctx->cfg.carrier_mode == GPSIQ_CARRIER_FLOAT) {
        k_slice_advance<<<C, 32, 0, st>>>(ctx->d_drift + 3 * (size_t) ctx->E * C, ctx->d_drift + (size_t) EC, ctx->d_adv, n_epochs, C);
        ctx->launches += 1;
        trace_mark(ctx, st, "k_slice_advance");
    } else {  // integer carrier: the batch's exact advance per slot (what another slice's owner folds into its state)
        k_int_carrier<<<C, 32, 0, st>>>(desc_dev, NULL, NULL, NULL, ctx->d_adv, 1, n_epochs, C, N);
        ctx->launches += 1;
        trace_mark(ctx, st, "k_int_carrier(adv)");
    }
    CU(cudaEventRecord(set.adv_done, st));
    ctx->last_epochs = n_epochs;
    CU(cudaGetLastError());
    return GPSIQ_OK;
}

// Phase 1b: everything that does NOT need the exact carrier phase: the code-NCO scan and the
// speculative carrier scans from the context's start-phase estimate (advanced afterwards).
// est: the start-phase estimate to speculate from -- the context's running one (advanced by the batch's closed-form
// advance afterwards: single-stream use and the phase API of time-sliced runs), or a set's own (pipelined submits).
static int enqueue_speculate(gpsiq_ctx* ctx, const gpsiq_chan_desc* desc_dev, int n_epochs, cudaStream_t st,
                             double* est = NULL) {
    const bool own_est = est != NULL;
    if (!est) est = ctx->d_est_state;
    cudaStream_t aux = ctx->aux2_stream;  // (aux_stream may be busy with the tile prologues of the batch being rendered)
    cudaEvent_t fork = ctx->ev_fork;
    ScanSet& sset = ctx->sets[ctx->set_spec];
    if (sset.phase != 1) return fail(ctx, GPSIQ_ERR_ARG, "speculate: the next batch in line has not been prepared", cudaSuccess);
    use_set(ctx, ctx->set_spec);
    const int C = ctx->C, N = ctx->N, T = ctx->T, ntiles = ctx->ntiles;
    const int EC = n_epochs * C;
    // the code-NCO scan (only k_synth_lanes needs one: k_synth_line's code anchors are closed form) does not depend on
    // the carrier chain: it runs beside it on the aux stream
    if (!ctx->use_line) {
        CU(cudaEventRecord(fork, st));
        CU(cudaStreamWaitEvent(aux, fork, 0));
        k_scan_code<<<(EC + 63) / 64, 64, 0, aux>>>(desc_dev, ctx->d_code_ck, ctx->d_wrap_ck, NULL, NULL, EC, C, N, T, ntiles);
        CU(cudaEventRecord(ctx->ev_code2, aux));
        trace_mark(ctx, aux, "k_scan_code");
        ctx->launches += 1;
    }
    if (ctx->cfg.carrier_mode == GPSIQ_CARRIER_FLOAT && ctx->cfg.reserved[0] == 0) {
        const size_t ECmax = (size_t) ctx->E * C;
        double* eadv = ctx->d_drift;
        double* ereset = ctx->d_drift + (size_t) EC;      // k_prepare wrote it at [gridDim.x + ec] with gridDim.x = EC
        double* est_epoch = ctx->d_drift + 2 * ECmax;
        trace_mark(ctx, st, "(speculate begin)");
        const bool free_run = ctx->free_running && ctx->slice_spec && !own_est;
        if (free_run) {
            k_est_open_loop<<<1, 32, 0, st>>>(est, ctx->d_est_meas, ctx->d_est_ccum, sset.d_ccum_used, C);
            ctx->launches += 1;
        }
        k_epoch_estimates<<<C, 32, 0, st>>>(ctx->d_drift + 3 * ECmax, ereset, est, est_epoch, n_epochs, C);
        trace_mark(ctx, st, "k_epoch_estimates");
        const int chains = EC * ctx->J;
        k_carr_speculate<<<(chains + 127) / 128, 128, 0, st>>>(desc_dev, eadv, est_epoch, ctx->d_carr_ck,
                                                         ctx->ck_plane, ctx->d_spec, n_epochs, C, N, T, ntiles, ctx->G,
                                                         ctx->J);
        trace_mark(ctx, st, "k_carr_speculate");
        CU(cudaEventRecord(sset.spec_done, st));
        k_carr_stitch<<<(EC * 2 + 127) / 128, 128, 0, st>>>(desc_dev, est_epoch, ctx->d_spec, ctx->d_carr_ck,
                                                      ctx->ck_plane, ctx->d_cinfo, ctx->d_specE, n_epochs, C, N, T,
                                                      ntiles, ctx->G, ctx->J);
        trace_mark(ctx, st, "k_carr_stitch");
        {
            const int ngroups = (n_epochs + GROUP_EPOCHS - 1) / GROUP_EPOCHS;
            k_carr_group<<<ngroups * C * 2, 32, 0, st>>>(desc_dev, ctx->d_specE, est_epoch,
                                                                   ctx->d_carr_ck, ctx->ck_plane, ctx->d_info, ECmax,
                                                                   ctx->d_traceG, ctx->d_specG, ctx->d_tieG, ctx->d_fallbacks,
                                                                   n_epochs, C, N, T, ntiles);
        }
        trace_mark(ctx, st, "k_carr_group");
        if (ctx->slice_spec) {
            k_carr_slice<<<C * 2, 32, 0, st>>>(desc_dev, ctx->d_specE, ctx->d_specG, ctx->d_tieG, est_epoch, ctx->d_startS,
                                               ctx->d_specS, ctx->d_tieS, n_epochs, C, N);
            trace_mark(ctx, st, "k_carr_slice");
            ctx->launches += 1;
        }
        if (free_run) { k_est_from_slice<<<1, 32, 0, st>>>(ctx->d_est_state, ctx->d_specS, ctx->d_startS, ctx->d_adv, C); ctx->launches += 1; }
        else if (!own_est) { k_est_fold<<<1, 32, 0, st>>>(ctx->d_est_state, ctx->d_adv, C); ctx->launches += 1; }
        ctx->launches += 4;
    }
    CU(cudaEventRecord(sset.spec_all, st));
    sset.phase = 2;
    ctx->set_spec = (ctx->set_spec + 1) % NSETS;
    CU(cudaGetLastError());
    return GPSIQ_OK;
}

static int enqueue_anchor(gpsiq_ctx* ctx, ScanSet& set, cudaStream_t st, cudaStream_t st_patch);

// Phase 1c: the serial part -- chain the exact carrier phase through the batch (advances the carrier
// state) and re-anchor the estimate on it.
static int enqueue_chain(gpsiq_ctx* ctx, const gpsiq_chan_desc* desc_dev, int n_epochs, cudaStream_t st) {
    const int C = ctx->C, N = ctx->N, T = ctx->T, ntiles = ctx->ntiles;
    if (ctx->sets[ctx->set_wr].phase != 2) return fail(ctx, GPSIQ_ERR_ARG, "chain: the next batch in line has not been speculated", cudaSuccess);
    use_set(ctx, ctx->set_wr);
    ScanSet& wset = ctx->sets[ctx->set_wr];
    bool float_chain = false;
    CU(cudaStreamWaitEvent(st, ctx->sets[ctx->set_wr].spec_all, 0));  // (a no-op when the speculation ran on this stream)
    CU(cudaStreamWaitEvent(st, ctx->ev_final, 0));  // the carrier state: after the previous batch's chain, whatever its stream
    if (ctx->cfg.carrier_mode == GPSIQ_CARRIER_FLOAT && ctx->cfg.reserved[0] == 0) {
        // one kernel on the path of the hand-off: it also leaves the batch's exact start / end phases in the set (and
        // re-anchors the estimate), so that nothing else has to be enqueued between the chain and the caller's send
        k_carr_final<<<C, 32, 0, st>>>(desc_dev, ctx->d_specE, ctx->d_specG, ctx->d_carr_ck, ctx->ck_plane,
                                       ctx->d_info, (size_t) ctx->E * C, ctx->d_traceG, ctx->d_carr_state, ctx->d_carr_trace,
                                       ctx->d_ginfo, ctx->d_fallbacks, ctx->d_tieG, ctx->slice_spec ? ctx->d_specS : NULL,
                                       ctx->d_tieS, ctx->d_sres, ctx->d_slice_stats,
                                       ctx->d_start0, wset.d_exact_end,
                                       (ctx->chain_keeps_estimate || ctx->free_running) ? NULL : ctx->d_est_state,
                                       ctx->handoff, ctx->d_startS, wset.d_ccum_used,
                                       (ctx->free_running && ctx->slice_spec) ? ctx->d_est_meas : NULL, n_epochs, C, N, T, ntiles);
        memset(&ctx->handoff, 0, sizeof ctx->handoff);
        float_chain = true;
    } else if (ctx->cfg.carrier_mode == GPSIQ_CARRIER_INT32) {  // closed form: one prefix sum over the epochs
        k_int_carrier<<<C, 32, 0, st>>>(desc_dev, ctx->d_ustart, ctx->d_carr_state, ctx->d_carr_trace, NULL, 0, n_epochs, C, N);
    } else {  // the serial float scan (cfg.reserved[0] = 1, cross-check)
        k_scan_carrier<<<C, 32, 0, st>>>(desc_dev, ctx->d_carr_ck + 6 * ctx->ck_plane, ctx->d_carr_state,
                                         ctx->d_carr_trace, ctx->d_info + 2 * (size_t) ctx->E * C, ctx->d_ginfo, GROUP_EPOCHS,
                                         n_epochs, C, N, T, ntiles, ctx->cfg.carrier_mode);
    }
    ctx->launches += 1;
    trace_mark(ctx, st, "k_carr_final");
    ScanSet& set = ctx->sets[ctx->set_wr];
    if (!float_chain) {
        if (!ctx->chain_keeps_estimate)  // re-anchor the estimate on the exact phase (single-stream use)
            CU(cudaMemcpyAsync(ctx->d_est_state, ctx->d_carr_state, C * sizeof(double), cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(set.d_exact_end, ctx->d_carr_state, C * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    CU(cudaEventRecord(ctx->ev_final, st));
    // everything below is off the hand-off path: on the set's own stream (time-sliced runs: the caller's next
    // operation on `st` is the send of the end phases)
    cudaStream_t st2 = (ctx->use_line || float_chain) ? set.stream : st;
    if (st2 != st) {
        CU(cudaEventRecord(ctx->ev_chain, st));
        CU(cudaStreamWaitEvent(st2, ctx->ev_chain, 0));
    }
    if (float_chain) {
        if (ctx->slice_spec) {
            const int ngroups = (n_epochs + GROUP_EPOCHS - 1) / GROUP_EPOCHS;
            k_carr_final_groups<<<ngroups * C, 32, 0, st2>>>(desc_dev, ctx->d_specE, ctx->d_specG, ctx->d_carr_ck, ctx->ck_plane,
                                                            ctx->d_info, (size_t) ctx->E * C, ctx->d_traceG, ctx->d_carr_trace,
                                                            ctx->d_ginfo, ctx->d_fallbacks, ctx->d_tieG, ctx->d_startS, ctx->d_sres,
                                                            ctx->d_start0, set.d_exact_end, ctx->d_err, n_epochs, C, N, T, ntiles);
            trace_mark(ctx, st2, "k_carr_final_groups");
            ctx->launches += 1;
        }
        k_bias_update<<<1, 32, 0, st2>>>(ctx->d_adv, ctx->d_start0, set.d_exact_end, ctx->d_bias_rate, set.d_rate_used, n_epochs, C);
        ctx->launches += 1;
    }
    if (!ctx->use_line) CU(cudaStreamWaitEvent(st2, ctx->ev_code2, 0));  // scan_done covers the code scan on the side stream too
    CU(cudaEventRecord(set.scan_done, st2));   // (the batch's scan results are complete only now)
    set.phase = 3;
    ctx->set_wr = (ctx->set_wr + 1) % NSETS;
    ctx->set_pending++;
    CU(cudaGetLastError());
    // Line kernel: tile anchors, safety check and patch walk depend only on the scan results: right behind the chain, on
    // the set's own stream -- long before the batch is rendered, and beside (not in front of) whatever the caller puts
    // on `st` next (in time-sliced runs: the hand-off of the exact phases to the next GPU).
    if (ctx->use_line) {
        if (st2 != set.stream) CU(cudaStreamWaitEvent(set.stream, set.scan_done, 0));
        return enqueue_anchor(ctx, set, set.stream, set.stream);
    }
    return GPSIQ_OK;
}

static int enqueue_scan(gpsiq_ctx* ctx, const gpsiq_chan_desc* desc_dev, int n_epochs, cudaStream_t st) {
    int rc = enqueue_prepare(ctx, desc_dev, n_epochs, st);
    if (!rc) rc = enqueue_speculate(ctx, desc_dev, n_epochs, st);
    if (!rc) rc = enqueue_chain(ctx, desc_dev, n_epochs, st);
    return rc;
}

// Phase 2: the per-sample synthesis from the checkpoints of the last scan.
static CarrLookup make_lookup(const gpsiq_ctx* ctx) {
    CarrLookup L;
    L.ck = ctx->d_carr_ck; L.plane = ctx->ck_plane; L.ginfo = ctx->d_ginfo; L.infoG = ctx->d_info;
    L.info_plane = (size_t) ctx->E * ctx->C; L.cinfo = ctx->d_cinfo; L.spec = ctx->d_spec; L.G = ctx->G; L.J = ctx->J; L.GP = GROUP_EPOCHS;
    return L;
}

// Phase 1d (line kernel): tile anchors, the safety check and the literal-recurrence walk of the tiles it cannot clear,
// for the set the working pointers are on (enqueue_chain calls it).
static int enqueue_anchor(gpsiq_ctx* ctx, ScanSet& set, cudaStream_t st, cudaStream_t st_patch) {
    const gpsiq_chan_desc* desc_dev = set.desc;
    const int n_epochs = set.n_epochs;
    const int C = ctx->C, N = ctx->N, ntiles = ctx->ntiles;
    const int dbg = ctx->cfg.reserved[1];
    ulonglong2* anch = ctx->d_anch[ctx->set_cur];
    CU(cudaMemsetAsync(ctx->d_line_counters, 0, 4 * sizeof(int), st));
    const int warps = n_epochs * C;
    const int intc = ctx->cfg.carrier_mode == GPSIQ_CARRIER_INT32;
    const uint32_t* ustart = intc ? ctx->d_ustart : NULL;
    k_line_anchor<<<(warps + 3) / 4, 128, 0, st>>>(desc_dev, make_lookup(ctx), ustart, ctx->d_flags, ctx->d_flags + ctx->E,
                                                   anch, ctx->d_lrecs, ctx->d_hazlist, ctx->d_line_counters,
                                                   ctx->haz_cap, n_epochs, C, N, ntiles, dbg);
    trace_mark(ctx, st, "k_line_anchor");
    k_line_check<<<(warps + 127) / 128, 128, 0, st>>>(desc_dev, ctx->d_lrecs, ctx->d_elist, ctx->d_line_counters, warps, N,
                                                      intc, dbg);
    k_line_refine<<<(dbg & LN_DBG_FORCE_CHUNK) ? 1024 : 64, 128, 0, st>>>(
        desc_dev, ctx->d_lrecs, make_lookup(ctx), ustart, ctx->d_elist, ctx->d_flags + ctx->E, ctx->d_hazlist,
        ctx->d_line_counters, ctx->haz_cap, C, N, ntiles, dbg);
    trace_mark(ctx, st, "k_line_refine");
    // The (tile, slot) pairs the check could not clear (~1e-4 of them): exact code-NCO state from the epoch's start,
    // then the literal recurrence over the tile, compared with the anchors' lines -> patch list.  (If a list
    // overflows the epoch is flagged and k_synth_lanes, which runs last, re-renders it.)
    CU(cudaEventRecord(set.anch_ready, st));
    if (st_patch != st) CU(cudaStreamWaitEvent(st_patch, set.anch_ready, 0));
    k_line_patch<<<(dbg & LN_DBG_FORCE_TILE) ? 1024 : 8, 128, 0, st_patch>>>(
        desc_dev, ctx->d_lutp, make_lookup(ctx), ustart, anch, ctx->d_chips4, ctx->d_hazlist,
        ctx->d_line_counters, ctx->haz_cap, ctx->d_patches, ctx->patch_cap, ctx->d_flags + ctx->E, C, N, ntiles);
    trace_mark(ctx, st_patch, "k_line_patch");
    CU(cudaEventRecord(set.anchor_done, st_patch));
    set.anchored = 1;
    ctx->launches += 4;
    CU(cudaGetLastError());
    return GPSIQ_OK;
}

static int enqueue_render(gpsiq_ctx* ctx, int16_t* iq_dev, cudaStream_t st, int16_t* iq_host = NULL) {
    if (ctx->set_pending < 1 || ctx->sets[ctx->set_rd].phase != 3)
        return fail(ctx, GPSIQ_ERR_ARG, "nothing to render (scan phases of a batch must complete first)", cudaSuccess);
    ScanSet& set = ctx->sets[ctx->set_rd];
    CU(cudaStreamWaitEvent(st, set.scan_done, 0));
    use_set(ctx, ctx->set_rd);
    const gpsiq_chan_desc* desc_dev = set.desc;
    const int n_epochs = set.n_epochs;
    const int C = ctx->C, N = ctx->N, T = ctx->T, ntiles = ctx->ntiles;
    if (ctx->ev_count < TIMING_RING) CU(cudaEventRecord(ctx->ev[ctx->ev_count][1], st));
    trace_mark(ctx, st, "(render begin)");
    const int tile_groups = (ntiles + LANES_WARPS - 1) / LANES_WARPS;
    const size_t smem = (size_t) C * 512 * sizeof(int2) + (size_t) C * 33 * 4;
    // everything below needs the chain's result; the aux stream (already holding the code scan) joins here
    CU(cudaEventRecord(ctx->ev_chain, st));
    CU(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_chain, 0));
    if (ctx->use_line) {
        ulonglong2* anch = ctx->d_anch[ctx->set_cur];
        const int intc = ctx->cfg.carrier_mode == GPSIQ_CARRIER_INT32;
        if (!set.anchored) {  // phase API: anchors here, the patch walk on the side stream beside the sample kernel
            int rca = enqueue_anchor(ctx, set, st, ctx->aux_stream);
            if (rca) return rca;
        }
        CU(cudaStreamWaitEvent(st, set.anch_ready, 0));
        if (ctx->render_waits_spec && ctx->set_pending >= 2 && ctx->sets[(ctx->set_rd + 1) % NSETS].phase >= 2 &&
            ctx->cfg.reserved[0] == 0 && !intc) {
            // Another batch has been submitted ahead.  Its chunk speculation wants the whole GPU (one chain per
            // thread, as many resident as possible) while the sample kernel below is issue-bound and holds on to
            // the SMs it gets: let the speculation finish first (the anchor kernel above ran beside it); the rest
            // of that batch's carrier chain (few, latency-bound threads) then runs beside the sample kernel.
            // Time-sliced runs (GPSIQ_OPT_RENDER_AFTER_NEXT_CHAIN) wait for that batch's whole chain instead: its
            // exact step is a hop of the inter-GPU ring, and a hop beside the sample kernel is several times slower
            // (the chain kernel and the NCCL hand-off wait for SM resources) -- that delay multiplies by the ring
            // length, while waiting here costs this rank at most the ring's quiet length once.
            ScanSet& nxt = ctx->sets[(ctx->set_rd + 1) % NSETS];
            CU(cudaStreamWaitEvent(st, (ctx->render_after_next_chain && nxt.phase >= 3) ? nxt.scan_done : nxt.spec_done, 0));
        }
        // device-resident output: one launch; host output: sub-batches so that the copies overlap the rendering
        const int sub = iq_host ? 32 : n_epochs;
        int k = 0;
        for (int e0 = 0; e0 < n_epochs; e0 += sub, k++) {
            const int ne = n_epochs - e0 < sub ? n_epochs - e0 : sub;
            int grid = ne * ((ntiles + LN_UNIT - 1) / LN_UNIT);
            if (ctx->line_grid_cap > 0 && grid > ctx->line_grid_cap) grid = ctx->line_grid_cap;  // CTAs stride over the units
            const bool timed = (k == 0 && ctx->ev_count < TIMING_RING);
            if (timed) { CU(cudaEventRecord(ctx->ev[ctx->ev_count][3], st)); ctx->fixed_epochs = ne; }
            k_synth_line<<<grid, LN_THREADS, ln_smem_bytes(C), st>>>(
                desc_dev + (size_t) e0 * C, ctx->d_lutp + (size_t) e0 * C * 512, ctx->d_chips4,
                anch + (size_t) e0 * ntiles * C, ctx->d_flags + e0, ctx->d_flags + ctx->E + e0,
                iq_dev + (size_t) e0 * N * 2, ne, C, N, ntiles, intc, ctx->d_err);
            if (timed) CU(cudaEventRecord(ctx->ev[ctx->ev_count][4], st));
            trace_mark(ctx, st, "k_synth_line");
            ctx->last_ln.desc = desc_dev + (size_t) e0 * C; ctx->last_ln.iq = iq_dev + (size_t) e0 * N * 2;
            ctx->last_ln.ne = ne; ctx->last_ln.set = ctx->set_cur;
            ctx->last_ln.e0 = e0;
            if (k == 0) {
                CU(cudaStreamWaitEvent(st, set.anchor_done, 0));  // the patch list is complete (and the epoch flags final)
                // exact code-NCO states for the epochs k_synth_lanes has to render (normally none: every thread returns)
                k_scan_code<<<(n_epochs * C + 63) / 64, 64, 0, st>>>(desc_dev, ctx->d_code_ck, ctx->d_wrap_ck, ctx->d_flags,
                                                                      ctx->d_flags + ctx->E, n_epochs * C, C, N, T, ntiles);
                ctx->launches += 1;
            }
            k_line_apply<<<4, 128, 0, st>>>(ctx->d_patches, ctx->d_line_counters, ctx->patch_cap,
                                            reinterpret_cast<uint32_t*>(iq_dev), (unsigned long long) e0 * N,
                                            (unsigned long long) (e0 + ne) * N, e0 == 0 ? ctx->d_line_totals : NULL);
            // epochs outside the line kernel's contract (or whose hazard / patch lists overflowed)
            k_synth_lanes<<<ne * LANES_FLAGGED_CTAS, LANES_WARPS * 32, smem, st>>>(
                desc_dev, ctx->d_lut, ctx->d_code_ck, ctx->d_wrap_ck, make_lookup(ctx), ctx->d_ustart, ctx->d_ca, ctx->d_flags,
                ctx->d_flags + ctx->E, 1, iq_dev, e0, C, N, T, ntiles, tile_groups, LANES_FLAGGED_CTAS, ctx->cfg.carrier_mode);
            ctx->launches += 3;
            if (iq_host) {
                CU(cudaEventRecord(ctx->ev_F[k & 1], st));
                CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_F[k & 1], 0));
                CU(cudaMemcpyAsync(iq_host + (size_t) e0 * N * 2, iq_dev + (size_t) e0 * N * 2, (size_t) ne * N * 4,
                                   cudaMemcpyDeviceToHost, ctx->copy_stream));
            }
        }
    } else {
        CU(cudaEventRecord(ctx->ev_P[0], ctx->aux_stream));   // join: the code scan must be complete
        CU(cudaStreamWaitEvent(st, ctx->ev_P[0], 0));
        k_synth_lanes<<<n_epochs * tile_groups, LANES_WARPS * 32, smem, st>>>(
            desc_dev, ctx->d_lut, ctx->d_code_ck, ctx->d_wrap_ck, make_lookup(ctx), ctx->d_ustart, ctx->d_ca, ctx->d_flags,
            ctx->d_flags + ctx->E, 0, iq_dev, 0, C, N, T, ntiles, tile_groups, tile_groups, ctx->cfg.carrier_mode);
        ctx->launches += 1;
        if (iq_host)
            CU(cudaMemcpyAsync(iq_host, iq_dev, (size_t) n_epochs * N * 4, cudaMemcpyDeviceToHost, st));
    }
    if (ctx->ev_count < TIMING_RING) {
        CU(cudaEventRecord(ctx->ev[ctx->ev_count][2], st));
        ctx->ev_count++;
    }
    trace_mark(ctx, st, "(render end)");
    CU(cudaEventRecord(set.render_done, st));
    set.phase = 0;
    ctx->set_rd = (ctx->set_rd + 1) % NSETS;
    ctx->set_pending--;
    CU(cudaGetLastError());
    return GPSIQ_OK;
}

static int enqueue(gpsiq_ctx* ctx, const gpsiq_chan_desc* desc_dev, int n_epochs, int16_t* iq_dev, cudaStream_t st) {
    if (ctx->set_pending) return fail(ctx, GPSIQ_ERR_ARG, "batches submitted ahead are still in flight: fetch them first", cudaSuccess);
    int rc = enqueue_scan(ctx, desc_dev, n_epochs, st);
    if (rc) return rc;
    return enqueue_render(ctx, iq_dev, st);
}

static int check_device_error(gpsiq_ctx* ctx) {
    int h = 0;
    CU(cudaMemcpyAsync(&h, ctx->d_err, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (h) {
        CU(cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream));
        if (h & 0x78000000) {
            snprintf(ctx->err, sizeof ctx->err, "internal device error 0x%x (%s)", h,
                     (h & GPSIQ_DEVERR_SLICE) ? "the parallel group chain disagrees with the slice-level translation"
                     : (h & GPSIQ_DEVERR_HANDOFF) ? "the previous slice's carrier phases never arrived in the mailbox"
                                                  : "a table copy did not complete");
            return GPSIQ_ERR_CUDA;
        }
        snprintf(ctx->err, sizeof ctx->err, "descriptor %d (epoch*max_chan+slot) is out of contract", h - 1);
        return GPSIQ_ERR_ARG;
    }
    return GPSIQ_OK;
}

int gpsiq_synth(gpsiq_ctx* ctx, const gpsiq_chan_desc* desc, int n_epochs, int16_t* iq_out) {
    if (!ctx || !desc || n_epochs < 1) return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_synth: bad argument", cudaSuccess);
    if (n_epochs > ctx->E) return fail(ctx, GPSIQ_ERR_CAPACITY, "gpsiq_synth: n_epochs > max_epochs", cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaMemcpyAsync(ctx->d_desc, desc, (size_t) n_epochs * ctx->C * sizeof(gpsiq_chan_desc), cudaMemcpyHostToDevice,
                       ctx->stream));
    if (ctx->set_pending) return fail(ctx, GPSIQ_ERR_ARG, "batches submitted ahead are still in flight: fetch them first", cudaSuccess);
    int rc = enqueue_scan(ctx, ctx->d_desc, n_epochs, ctx->stream);
    if (!rc) rc = enqueue_render(ctx, ctx->d_iq, ctx->stream, iq_out);
    if (rc) return rc;
    rc = check_device_error(ctx);
    CU(cudaStreamSynchronize(ctx->copy_stream));
    return rc;
}

int gpsiq_synth_device(gpsiq_ctx* ctx, const gpsiq_chan_desc* desc_dev, int n_epochs, int16_t* iq_dev, void* stream) {
    if (!ctx || !desc_dev || !iq_dev || n_epochs < 1 || ((uintptr_t) iq_dev & 15))
        return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_synth_device: bad argument", cudaSuccess);
    if (n_epochs > ctx->E) return fail(ctx, GPSIQ_ERR_CAPACITY, "gpsiq_synth_device: n_epochs > max_epochs", cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    return enqueue(ctx, desc_dev, n_epochs, iq_dev, (cudaStream_t) stream);
}

// Streaming pair: submit scans a batch ahead on the context's own stream (into the free scan set),
// fetch renders the oldest submitted batch on the caller's stream.  With one batch of lookahead the
// serial carrier chain of batch k+1 overlaps the sample kernels of batch k.
// Start-phase estimate of a pipelined batch (estimates only ever affect speed: a poor one makes epochs fall back to
// the serial scan).  Nothing in flight: the exact carrier state.  Otherwise the scans of consecutive batches overlap,
// so the previous batch's exact end is not known yet: the estimate is the exact end of the batch BEFORE it (its chain
// was enqueued a whole batch earlier) advanced by the previous batch's closed-form advance -- or, at the start of a
// burst, the previous batch's own estimate advanced the same way.
static int pipelined_estimate(gpsiq_ctx* ctx, ScanSet& set, int idx, cudaStream_t ss) {
    const size_t bytes = ctx->C * sizeof(double);
    if (ctx->set_pending == 0) {
        CU(cudaStreamWaitEvent(ss, ctx->ev_final, 0));
        CU(cudaMemcpyAsync(set.d_est, ctx->d_carr_state, bytes, cudaMemcpyDeviceToDevice, ss));
    } else {
        ScanSet& prev = ctx->sets[(idx + NSETS - 1) % NSETS];
        ScanSet& prev2 = ctx->sets[(idx + NSETS - 2) % NSETS];
        if (prev.seq != set.seq - 1) return fail(ctx, GPSIQ_ERR_ARG, "internal: scan set ring out of order", cudaSuccess);
        if (NSETS >= 3 && prev2.seq == set.seq - 2 && prev2.phase != 1 && prev2.phase != 2) {
            CU(cudaStreamWaitEvent(ss, prev2.scan_done, 0));
            CU(cudaMemcpyAsync(set.d_est, prev2.d_exact_end, bytes, cudaMemcpyDeviceToDevice, ss));
        } else {
            CU(cudaStreamWaitEvent(ss, prev.est_done, 0));
            CU(cudaMemcpyAsync(set.d_est, prev.d_est, bytes, cudaMemcpyDeviceToDevice, ss));
        }
        CU(cudaStreamWaitEvent(ss, prev.adv_done, 0));
        k_est_fold<<<1, 32, 0, ss>>>(set.d_est, prev.d_adv, ctx->C);
        ctx->launches += 1;
    }
    CU(cudaEventRecord(set.est_done, ss));
    return GPSIQ_OK;
}

// desc_dev == NULL: the descriptors are already in the write set's own buffer (host submit)
static int submit_common(gpsiq_ctx* ctx, const gpsiq_chan_desc* desc_dev, int n_epochs, cudaStream_t after) {
    if (ctx->set_prep != ctx->set_wr || ctx->set_spec != ctx->set_wr)
        return fail(ctx, GPSIQ_ERR_ARG, "a batch begun through the phase calls is still waiting for its chain", cudaSuccess);
    const int idx = ctx->set_wr;
    ScanSet& set = ctx->sets[idx];
    cudaStream_t ss = set.stream;
    if (after) {  // the descriptors are produced on the caller's stream
        CU(cudaEventRecord(ctx->ev_fork2, after));
        CU(cudaStreamWaitEvent(ss, ctx->ev_fork2, 0));
    }
    int rc = begin_batch(ctx, ss);
    if (rc) return rc;
    if (desc_dev)
        CU(cudaMemcpyAsync(set.d_descbuf, desc_dev, (size_t) n_epochs * ctx->C * sizeof(gpsiq_chan_desc),
                           cudaMemcpyDeviceToDevice, ss));
    rc = enqueue_prepare(ctx, set.d_descbuf, n_epochs, ss, true);
    const bool spec = ctx->cfg.carrier_mode == GPSIQ_CARRIER_FLOAT && ctx->cfg.reserved[0] == 0;
    if (!rc && spec) rc = pipelined_estimate(ctx, set, idx, ss);
    if (!rc) rc = enqueue_speculate(ctx, set.d_descbuf, n_epochs, ss, spec ? set.d_est : NULL);
    if (!rc) rc = enqueue_chain(ctx, set.d_descbuf, n_epochs, ss);
    return rc;   // (the anchors follow the chain on the set's stream: enqueue_chain)
}

// Streaming pair: submit scans a batch ahead on one of the context's own streams (into the free scan set),
// fetch renders the oldest submitted batch on the caller's stream.  Up to NSETS - 1 batches may be scanned ahead of
// the one being rendered; their scans overlap each other and the sample kernels.
int gpsiq_submit_device(gpsiq_ctx* ctx, const gpsiq_chan_desc* desc_dev, int n_epochs, void* after_stream) {
    if (!ctx || !desc_dev || n_epochs < 1) return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_submit_device: bad argument", cudaSuccess);
    if (n_epochs > ctx->E) return fail(ctx, GPSIQ_ERR_CAPACITY, "gpsiq_submit_device: n_epochs > max_epochs", cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    return submit_common(ctx, desc_dev, n_epochs, (cudaStream_t) after_stream);
}

int gpsiq_fetch_device(gpsiq_ctx* ctx, int16_t* iq_dev, void* stream) {
    if (!ctx || !iq_dev || ((uintptr_t) iq_dev & 15)) return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_fetch_device: bad argument", cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    return enqueue_render(ctx, iq_dev, (cudaStream_t) stream);
}

// Host-buffer streaming pair: same pipeline, descriptors from / samples to HOST memory.
int gpsiq_submit(gpsiq_ctx* ctx, const gpsiq_chan_desc* desc, int n_epochs) {
    if (!ctx || !desc || n_epochs < 1) return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_submit: bad argument", cudaSuccess);
    if (n_epochs > ctx->E) return fail(ctx, GPSIQ_ERR_CAPACITY, "gpsiq_submit: n_epochs > max_epochs", cudaSuccess);
    if (ctx->set_prep != ctx->set_wr || ctx->set_spec != ctx->set_wr)
        return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_submit: a batch begun through the phase calls is still waiting for its chain", cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    const size_t bytes = (size_t) n_epochs * ctx->C * sizeof(gpsiq_chan_desc);
    const int w = ctx->set_wr;
    if (!ctx->h_stage[w]) CU(cudaHostAlloc(&ctx->h_stage[w], (size_t) ctx->E * ctx->C * sizeof(gpsiq_chan_desc), cudaHostAllocDefault));
    CU(cudaEventSynchronize(ctx->sets[w].scan_done));  // the staging buffer's previous upload has been consumed
    memcpy(ctx->h_stage[w], desc, bytes);
    // (the set's previous batch has been rendered: fetch is blocking, and a set is only rewritten after its fetch)
    if (ctx->sets[w].phase != 0) return fail(ctx, GPSIQ_ERR_CAPACITY, "gpsiq_submit: every scan set is in flight (fetch a batch first)", cudaSuccess);
    CU(cudaStreamWaitEvent(ctx->sets[w].stream, ctx->sets[w].render_done, 0));  // its previous batch no longer reads the buffer
    CU(cudaMemcpyAsync(ctx->sets[w].d_descbuf, ctx->h_stage[w], bytes, cudaMemcpyHostToDevice, ctx->sets[w].stream));
    return submit_common(ctx, NULL, n_epochs, NULL);
}

int gpsiq_fetch(gpsiq_ctx* ctx, int16_t* iq_out) {
    if (!ctx || !iq_out) return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_fetch: bad argument", cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    if (!ctx->d_iq2) CU(cudaMalloc(&ctx->d_iq2, (size_t) ctx->E * ctx->N * 4));
    int16_t* dev = (ctx->fetch_count++ & 1) ? ctx->d_iq2 : ctx->d_iq;
    int rc = enqueue_render(ctx, dev, ctx->stream, iq_out);
    if (rc) return rc;
    rc = check_device_error(ctx);
    CU(cudaStreamSynchronize(ctx->copy_stream));
    return rc;
}

// ---- one stream over several GPUs of this process ---------------------------------------------------------------
// The C-level multi-GPU driver ("one process, N devices"; the reference's epoch loop, plutogpssim.c:2655-2806,
// over N devices): consecutive batches of ONE stream go to consecutive devices, each through the ordinary pipelined
// submit of its context.  The only state that crosses a batch boundary is the carrier phase (plutogpssim.c:2741-2746):
// the exact chain of batch j starts from a copy of the carrier state batch j-1 left on ITS device, ordered by that
// context's chain event (stream waits work across devices); start-phase estimates are folded from the batch before
// last exactly as in one context.  No host synchronisation, no kernel for the hand-off: two small async copies.
#define MULTI_MAX_DEV 16
struct MultiSlot { ScanSet* set; int dev; long long seq; };
struct gpsiq_multi {
    int n;
    gpsiq_ctx* sub[MULTI_MAX_DEV];
    long long seq_submit, seq_begin, seq_end;
    MultiSlot ring[MULTI_MAX_DEV * NSETS];
    char err[256];
};

static int mfail(gpsiq_multi* m, int code, const char* what, gpsiq_ctx* c) {
    snprintf(m ? m->err : g_err, 256, "%.60s%s%.180s", what, c ? ": " : "", c ? c->err : "");
    return code;
}

int gpsiq_multi_create(gpsiq_multi** out, const gpsiq_config* cfg, int n_devices) {
    if (!out || !cfg || n_devices < 1 || n_devices > MULTI_MAX_DEV) return fail(NULL, GPSIQ_ERR_ARG, "gpsiq_multi_create: bad argument", cudaSuccess);
    *out = NULL;
    gpsiq_multi* m = (gpsiq_multi*) calloc(1, sizeof *m);
    if (!m) return fail(NULL, GPSIQ_ERR_NOMEM, "gpsiq_multi_create: out of host memory", cudaSuccess);
    m->n = n_devices;
    for (int i = 0; i < n_devices; i++) {
        gpsiq_config c = *cfg;
        c.device = cfg->device + i;
        const int rc = gpsiq_create(&m->sub[i], &c);
        if (rc != GPSIQ_OK) {  // (the message is in the global error text)
            for (int k = 0; k < i; k++) gpsiq_destroy(m->sub[k]);
            free(m);
            return rc;
        }
    }
    for (int i = 0; i < n_devices; i++)  // direct peer copies where the topology allows them (else the driver stages them)
        for (int k = 0; k < n_devices; k++) {
            int can = 0;
            if (i != k && cudaDeviceCanAccessPeer(&can, cfg->device + i, cfg->device + k) == cudaSuccess && can) {
                cudaSetDevice(cfg->device + i);
                cudaDeviceEnablePeerAccess(cfg->device + k, 0);
            }
        }
    cudaGetLastError();
    *out = m;
    return GPSIQ_OK;
}

void gpsiq_multi_destroy(gpsiq_multi* m) {
    if (!m) return;
    for (int i = 0; i < m->n; i++) gpsiq_destroy(m->sub[i]);
    free(m);
}

const char* gpsiq_multi_last_error(const gpsiq_multi* m) { return m ? m->err : g_err; }
int gpsiq_multi_devices(const gpsiq_multi* m) { return m ? m->n : 0; }
int64_t gpsiq_multi_launch_count(const gpsiq_multi* m) {
    int64_t t = 0;
    for (int i = 0; m && i < m->n; i++) t += m->sub[i]->launches;
    return t;
}

#define MCU(call)                                                                             \
    do {                                                                                      \
        cudaError_t ce_ = (call);                                                             \
        if (ce_ != cudaSuccess) { fail(c, GPSIQ_ERR_CUDA, #call, ce_); return mfail(m, GPSIQ_ERR_CUDA, "gpsiq_multi", c); } \
    } while (0)

int gpsiq_multi_submit(gpsiq_multi* m, const gpsiq_chan_desc* desc, int n_epochs) {
    if (!m || !desc || n_epochs < 1) return mfail(m, GPSIQ_ERR_ARG, "gpsiq_multi_submit: bad argument", NULL);
    const long long j = m->seq_submit;
    const int n = m->n, R = n * NSETS;
    gpsiq_ctx* c = m->sub[j % n];
    if (n_epochs > c->E) return mfail(m, GPSIQ_ERR_CAPACITY, "gpsiq_multi_submit: n_epochs > max_epochs", NULL);
    if (c->sets[c->set_wr].phase != 0 || c->set_prep != c->set_wr || c->set_spec != c->set_wr)
        return mfail(m, GPSIQ_ERR_CAPACITY, "gpsiq_multi_submit: every scan set of the next device is in flight (fetch a batch first)", NULL);
    MCU(cudaSetDevice(c->cfg.device));
    const int w = c->set_wr;
    ScanSet& set = c->sets[w];
    cudaStream_t ss = set.stream;
    const size_t bytes = (size_t) n_epochs * c->C * sizeof(gpsiq_chan_desc), cbytes = c->C * sizeof(double);
    if (!c->h_stage[w]) MCU(cudaHostAlloc(&c->h_stage[w], (size_t) c->E * c->C * sizeof(gpsiq_chan_desc), cudaHostAllocDefault));
    MCU(cudaEventSynchronize(set.scan_done));       // the staging buffer's previous upload has been consumed
    memcpy(c->h_stage[w], desc, bytes);
    MCU(cudaStreamWaitEvent(ss, set.render_done, 0));
    MCU(cudaMemcpyAsync(set.d_descbuf, c->h_stage[w], bytes, cudaMemcpyHostToDevice, ss));
    int rc = begin_batch(c, ss);
    if (!rc) rc = enqueue_prepare(c, set.d_descbuf, n_epochs, ss, true);
    if (rc) return mfail(m, rc, "gpsiq_multi_submit", c);
    const bool spec = c->cfg.carrier_mode == GPSIQ_CARRIER_FLOAT && c->cfg.reserved[0] == 0;
    const MultiSlot* P = (j >= 1 && m->ring[(j - 1) % R].seq == j - 1) ? &m->ring[(j - 1) % R] : NULL;
    const MultiSlot* P2 = (j >= 2 && m->ring[(j - 2) % R].seq == j - 2) ? &m->ring[(j - 2) % R] : NULL;
    gpsiq_ctx* pc = (j >= 1) ? m->sub[(j - 1) % n] : NULL;   // the context holding the carrier state before this batch
    if (spec) {
        if (m->seq_submit == m->seq_begin || !P) {  // nothing in flight: the exact state
            gpsiq_ctx* src = pc ? pc : c;
            MCU(cudaStreamWaitEvent(ss, src->ev_final, 0));
            MCU(cudaMemcpyAsync(set.d_est, src->d_carr_state, cbytes, cudaMemcpyDefault, ss));
        } else {
            if (P2) {
                MCU(cudaStreamWaitEvent(ss, P2->set->scan_done, 0));
                MCU(cudaMemcpyAsync(set.d_est, P2->set->d_exact_end, cbytes, cudaMemcpyDefault, ss));
            } else {
                MCU(cudaStreamWaitEvent(ss, P->set->est_done, 0));
                MCU(cudaMemcpyAsync(set.d_est, P->set->d_est, cbytes, cudaMemcpyDefault, ss));
            }
            MCU(cudaStreamWaitEvent(ss, P->set->adv_done, 0));
            MCU(cudaMemcpyAsync(c->d_adv_prev, P->set->d_adv, 2 * cbytes, cudaMemcpyDefault, ss));
            k_est_fold<<<1, 32, 0, ss>>>(set.d_est, c->d_adv_prev, c->C);
            c->launches += 1;
        }
        MCU(cudaEventRecord(set.est_done, ss));
    }
    rc = enqueue_speculate(c, set.d_descbuf, n_epochs, ss, spec ? set.d_est : NULL);
    if (!rc && pc && pc != c) {  // hand-off: the carrier state the previous batch's chain left on its device
        MCU(cudaStreamWaitEvent(ss, c->ev_final, 0));    // (this device's own older chain has consumed its state)
        MCU(cudaStreamWaitEvent(ss, pc->ev_final, 0));
        MCU(cudaMemcpyAsync(c->d_carr_state, pc->d_carr_state, cbytes, cudaMemcpyDefault, ss));
    }
    if (!rc) rc = enqueue_chain(c, set.d_descbuf, n_epochs, ss);
    if (rc) return mfail(m, rc, "gpsiq_multi_submit", c);
    MultiSlot& slot = m->ring[j % R];
    slot.set = &set; slot.dev = (int) (j % n); slot.seq = j;
    m->seq_submit++;
    return GPSIQ_OK;
}

// Render the oldest submitted batch into iq_out (HOST memory; pinned memory keeps the copies asynchronous) without
// waiting; gpsiq_multi_fetch_end waits for the oldest begun batch.  At most two begun batches per device.
int gpsiq_multi_fetch_begin(gpsiq_multi* m, int16_t* iq_out) {
    if (!m || !iq_out) return mfail(m, GPSIQ_ERR_ARG, "gpsiq_multi_fetch_begin: bad argument", NULL);
    if (m->seq_begin >= m->seq_submit) return mfail(m, GPSIQ_ERR_ARG, "gpsiq_multi_fetch_begin: nothing submitted", NULL);
    if (m->seq_begin - m->seq_end >= 2LL * m->n) return mfail(m, GPSIQ_ERR_CAPACITY, "gpsiq_multi_fetch_begin: two batches per device already begun", NULL);
    gpsiq_ctx* c = m->sub[m->seq_begin % m->n];
    MCU(cudaSetDevice(c->cfg.device));
    if (!c->d_iq2) MCU(cudaMalloc(&c->d_iq2, (size_t) c->E * c->N * 4));
    int16_t* dev = (c->fetch_count++ & 1) ? c->d_iq2 : c->d_iq;
    const int rc = enqueue_render(c, dev, c->stream, iq_out);
    if (rc) return mfail(m, rc, "gpsiq_multi_fetch_begin", c);
    m->seq_begin++;
    return GPSIQ_OK;
}

int gpsiq_multi_fetch_end(gpsiq_multi* m) {
    if (!m || m->seq_end >= m->seq_begin) return mfail(m, GPSIQ_ERR_ARG, "gpsiq_multi_fetch_end: nothing begun", NULL);
    gpsiq_ctx* c = m->sub[m->seq_end % m->n];
    MCU(cudaSetDevice(c->cfg.device));
    const int rc = check_device_error(c);
    MCU(cudaStreamSynchronize(c->copy_stream));
    m->seq_end++;
    if (rc) return mfail(m, rc, "gpsiq_multi_fetch_end", c);
    return GPSIQ_OK;
}

int gpsiq_multi_fetch(gpsiq_multi* m, int16_t* iq_out) {
    const int rc = gpsiq_multi_fetch_begin(m, iq_out);
    return rc ? rc : gpsiq_multi_fetch_end(m);
}
#undef MCU

int gpsiq_scan_device(gpsiq_ctx* ctx, const gpsiq_chan_desc* desc_dev, int n_epochs, void* stream) {
    if (!ctx || !desc_dev || n_epochs < 1) return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_scan_device: bad argument", cudaSuccess);
    if (n_epochs > ctx->E) return fail(ctx, GPSIQ_ERR_CAPACITY, "gpsiq_scan_device: n_epochs > max_epochs", cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    return enqueue_scan(ctx, desc_dev, n_epochs, (cudaStream_t) stream);
}

int gpsiq_prepare_device(gpsiq_ctx* ctx, const gpsiq_chan_desc* desc_dev, int n_epochs, double* advance_dev, void* stream) {
    if (!ctx || !desc_dev || n_epochs < 1) return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_prepare_device: bad argument", cudaSuccess);
    if (n_epochs > ctx->E) return fail(ctx, GPSIQ_ERR_CAPACITY, "gpsiq_prepare_device: n_epochs > max_epochs", cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    int rc = enqueue_prepare(ctx, desc_dev, n_epochs, (cudaStream_t) stream);
    if (rc) return rc;
    if (advance_dev)
        CU(cudaMemcpyAsync(advance_dev, ctx->d_adv, 2 * ctx->C * sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t) stream));
    return GPSIQ_OK;
}

int gpsiq_speculate_device(gpsiq_ctx* ctx, const gpsiq_chan_desc* desc_dev, int n_epochs, void* stream) {
    if (!ctx || !desc_dev || n_epochs != ctx->sets[ctx->set_spec].n_epochs || ctx->sets[ctx->set_spec].phase != 1)
        return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_speculate_device: must follow gpsiq_prepare_device of the same batch", cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    return enqueue_speculate(ctx, desc_dev, n_epochs, (cudaStream_t) stream);
}

int gpsiq_chain_device(gpsiq_ctx* ctx, const gpsiq_chan_desc* desc_dev, int n_epochs, void* stream) {
    if (!ctx || !desc_dev || n_epochs != ctx->sets[ctx->set_wr].n_epochs || ctx->sets[ctx->set_wr].phase != 2)
        return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_chain_device: must follow gpsiq_speculate_device of the same batch", cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    return enqueue_chain(ctx, desc_dev, n_epochs, (cudaStream_t) stream);
}

int gpsiq_estimate_fold_device(gpsiq_ctx* ctx, const double* advance_dev, void* stream) {
    if (!ctx || !advance_dev) return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_estimate_fold_device: bad argument", cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    k_est_fold<<<1, 32, 0, (cudaStream_t) stream>>>(ctx->d_est_state, advance_dev, ctx->C);
    ctx->launches += 1;
    CU(cudaGetLastError());
    return GPSIQ_OK;
}

int gpsiq_carrier_fold_device(gpsiq_ctx* ctx, const double* advance_dev, void* stream) {
    if (!ctx || !advance_dev) return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_carrier_fold_device: bad argument", cudaSuccess);
    if (ctx->cfg.carrier_mode != GPSIQ_CARRIER_INT32)
        return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_carrier_fold_device: only the integer carrier has a closed-form advance", cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    k_int_fold<<<1, 32, 0, (cudaStream_t) stream>>>(ctx->d_carr_state, advance_dev, ctx->C);
    ctx->launches += 1;
    CU(cudaGetLastError());
    return GPSIQ_OK;
}

int gpsiq_estimate_from_device(gpsiq_ctx* ctx, const double* src_dev, void* stream) {
    if (!ctx || !src_dev) return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_estimate_from_device: bad argument", cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaMemcpyAsync(ctx->d_est_state, src_dev, ctx->C * sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t) stream));
    return GPSIQ_OK;
}

int gpsiq_estimate_anchor_device(gpsiq_ctx* ctx, void* stream) {
    if (!ctx) return GPSIQ_ERR_ARG;
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaMemcpyAsync(ctx->d_est_state, ctx->d_carr_state, ctx->C * sizeof(double), cudaMemcpyDeviceToDevice,
                       (cudaStream_t) stream));
    return GPSIQ_OK;
}

int gpsiq_set_option(gpsiq_ctx* ctx, int option, int value) {
    if (!ctx) return GPSIQ_ERR_ARG;
    if (option == GPSIQ_OPT_CHAIN_KEEPS_ESTIMATE) { ctx->chain_keeps_estimate = value != 0; return GPSIQ_OK; }
    if (option == GPSIQ_OPT_RENDER_AFTER_NEXT_CHAIN) { ctx->render_after_next_chain = value != 0; return GPSIQ_OK; }
    if (option == GPSIQ_OPT_FREE_RUNNING_ESTIMATE) { ctx->free_running = value != 0; return GPSIQ_OK; }
    if (option == GPSIQ_OPT_LINE_GRID_CAP) { ctx->line_grid_cap = value > 0 ? value : 0; return GPSIQ_OK; }
    return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_set_option: unknown option", cudaSuccess);
}

int gpsiq_estimate_to_device(gpsiq_ctx* ctx, double* dst_dev, void* stream) {
    if (!ctx || !dst_dev) return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_estimate_to_device: bad argument", cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaMemcpyAsync(dst_dev, ctx->d_est_state, ctx->C * sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t) stream));
    return GPSIQ_OK;
}

int gpsiq_estimate_correct_device(gpsiq_ctx* ctx, const double* exact_old_dev, const double* est_old_dev, double gain,
                                  void* stream) {
    if (!ctx || !exact_old_dev || !est_old_dev) return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_estimate_correct_device: bad argument", cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    k_est_correct<<<1, 32, 0, (cudaStream_t) stream>>>(ctx->d_est_state, exact_old_dev, est_old_dev, gain, ctx->C);
    ctx->launches += 1;
    CU(cudaGetLastError());
    return GPSIQ_OK;
}

int gpsiq_render_device(gpsiq_ctx* ctx, const gpsiq_chan_desc* desc_dev, int n_epochs, int16_t* iq_dev, void* stream) {
    (void) desc_dev;
    if (!ctx || !iq_dev || n_epochs < 1 || ((uintptr_t) iq_dev & 15) || ctx->set_pending < 1 ||
        n_epochs != ctx->sets[ctx->set_rd].n_epochs)
        return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_render_device: bad argument (renders the oldest batch whose scan phases are complete)",
                    cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    return enqueue_render(ctx, iq_dev, (cudaStream_t) stream);
}

int gpsiq_carrier_to_device(gpsiq_ctx* ctx, double* dst_dev, void* stream) {
    if (!ctx || !dst_dev) return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_carrier_to_device: bad argument", cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaMemcpyAsync(dst_dev, ctx->d_carr_state, ctx->C * sizeof(double), cudaMemcpyDeviceToDevice,
                       (cudaStream_t) stream));
    return GPSIQ_OK;
}

int gpsiq_carrier_from_device(gpsiq_ctx* ctx, const double* src_dev, void* stream) {
    if (!ctx || !src_dev) return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_carrier_from_device: bad argument", cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaMemcpyAsync(ctx->d_carr_state, src_dev, ctx->C * sizeof(double), cudaMemcpyDeviceToDevice,
                       (cudaStream_t) stream));
    return GPSIQ_OK;  // the estimate is NOT touched: a speculation from it may already be in flight
}

// ---- mailbox hand-off ------------------------------------------------------------------
// The carrier phases at a slice boundary (plutogpssim.c:2741-2746 chains them across the whole stream) are the only
// data a time-sliced run moves between GPUs: max_chan doubles per hop, latency-bound.  A hop through NCCL needs a
// kernel on both GPUs, which on a GPU saturated by the sample kernel waits for SM resources (measured: 23 us idle,
// ~0.36 ms busy).  The mailbox path uses none: the sender's copy engine writes the state into the receiver's
// mailbox (peer memory over NVLink) and a stream memory operation then writes a sequence number; the receiver's
// stream waits on that number with a stream memory operation.
#define MBOX_SLOT 1024
#define MBOX_FLAG 2048
#define MBOX_COUNTER 3072
#define MBOX_BYTES 4096
typedef CUresult (*mbox_fn64)(CUstream, CUdeviceptr, cuuint64_t, unsigned int);

int gpsiq_mailbox_create(gpsiq_ctx* ctx, void* handle_out) {
    if (!ctx || !handle_out) return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_mailbox_create: bad argument", cudaSuccess);
    if ((size_t) ctx->C * sizeof(double) > MBOX_SLOT) return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_mailbox_create: too many channels", cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    if (!ctx->d_mbox) {
        CU(cudaMalloc(&ctx->d_mbox, MBOX_BYTES));
        CU(cudaMemset(ctx->d_mbox, 0, MBOX_BYTES));
        cudaDriverEntryPointQueryResult q;
        CU(cudaGetDriverEntryPoint("cuStreamWriteValue64", &ctx->fn_write64, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess) ctx->fn_write64 = NULL;
        CU(cudaGetDriverEntryPoint("cuStreamWaitValue64", &ctx->fn_wait64, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess) ctx->fn_wait64 = NULL;
        if (!ctx->fn_write64 || !ctx->fn_wait64)
            return fail(ctx, GPSIQ_ERR_CUDA, "gpsiq_mailbox_create: no 64-bit stream memory operations in this driver", cudaSuccess);
        int can = 0;
        cudaDeviceGetAttribute(&can, cudaDevAttrCanFlushRemoteWrites, ctx->cfg.device);
        ctx->mbox_flush = can;
        CU(cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, ctx->d_mbox));
    memcpy(handle_out, &h, sizeof h);  // 64 bytes
    return GPSIQ_OK;
}

int gpsiq_mailbox_open(gpsiq_ctx* ctx, const void* handle, int peer_device) {
    if (!ctx || !handle || !ctx->d_mbox) return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_mailbox_open: bad argument / no own mailbox", cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    if (peer_device != ctx->cfg.device) {
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, ctx->cfg.device, peer_device));
        if (!can) return fail(ctx, GPSIQ_ERR_CUDA, "gpsiq_mailbox_open: no peer access to that device", cudaSuccess);
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    void* p = NULL;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->d_mbox_peer = (unsigned char*) p;
    return GPSIQ_OK;
}

// carrier state -> slot (seq & 1) of the next rank's mailbox, then its flag <- seq (both in stream order)
int gpsiq_mailbox_send(gpsiq_ctx* ctx, uint64_t seq, void* stream) {
    if (!ctx || !ctx->d_mbox_peer) return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_mailbox_send: no peer mailbox open", cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaMemcpyAsync(ctx->d_mbox_peer + (seq & 1) * MBOX_SLOT, ctx->d_carr_state, ctx->C * sizeof(double),
                       cudaMemcpyDefault, (cudaStream_t) stream));
    const CUresult r = ((mbox_fn64) ctx->fn_write64)((CUstream) stream, (CUdeviceptr) (ctx->d_mbox_peer + MBOX_FLAG), seq, 0);
    if (r != CUDA_SUCCESS) return fail(ctx, GPSIQ_ERR_CUDA, "gpsiq_mailbox_send: cuStreamWriteValue64 failed", cudaSuccess);
    trace_mark(ctx, (cudaStream_t) stream, "mbox_send(flag written)");
    return GPSIQ_OK;
}

// The hand-off fused into the chain kernel: [wait for message recv_seq in the own mailbox] -> chain -> [end phases as
// message send_seq into the next GPU's mailbox].  recv_seq / send_seq 0: that side is not wanted.
int gpsiq_chain_handoff_device(gpsiq_ctx* ctx, const gpsiq_chan_desc* desc_dev, int n_epochs, uint64_t recv_seq,
                               uint64_t send_seq, double* start_copy_dev, void* stream) {
    if (!ctx || !desc_dev || n_epochs < 1 || n_epochs > ctx->E) return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_chain_handoff_device: bad argument", cudaSuccess);
    if (ctx->cfg.carrier_mode != GPSIQ_CARRIER_FLOAT || ctx->cfg.reserved[0] != 0)
        return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_chain_handoff_device: float carrier with the parallel scan only", cudaSuccess);
    if ((recv_seq && !ctx->d_mbox) || (send_seq && !ctx->d_mbox_peer))
        return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_chain_handoff_device: no mailbox (gpsiq_mailbox_create / _open first)", cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    Handoff h;
    memset(&h, 0, sizeof h);
    if (recv_seq) {
        h.in_flag = (const unsigned long long*) (ctx->d_mbox + MBOX_FLAG);
        h.in_seq = recv_seq;
        h.in_slot = (const double*) (ctx->d_mbox + (recv_seq & 1) * MBOX_SLOT);
    }
    if (send_seq) {
        h.out_slot = (double*) (ctx->d_mbox_peer + (send_seq & 1) * MBOX_SLOT);
        h.out_flag = (unsigned long long*) (ctx->d_mbox_peer + MBOX_FLAG);
        h.out_seq = send_seq;
        h.counter = (unsigned int*) (ctx->d_mbox + MBOX_COUNTER);
    }
    h.start_copy = start_copy_dev;
    h.err = ctx->d_err;
    ctx->handoff = h;
    const int rc = enqueue_chain(ctx, desc_dev, n_epochs, (cudaStream_t) stream);
    memset(&ctx->handoff, 0, sizeof ctx->handoff);
    return rc;
}

// wait until the own mailbox's flag >= seq, then slot (seq & 1) -> carrier state (the estimate is not touched)
int gpsiq_mailbox_recv(gpsiq_ctx* ctx, uint64_t seq, void* stream) {
    if (!ctx || !ctx->d_mbox) return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_mailbox_recv: no mailbox", cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    trace_mark(ctx, (cudaStream_t) stream, "mbox_recv(wait enqueued behind this)");
    const unsigned int flags = CU_STREAM_WAIT_VALUE_GEQ | (ctx->mbox_flush ? CU_STREAM_WAIT_VALUE_FLUSH : 0);
    const CUresult r = ((mbox_fn64) ctx->fn_wait64)((CUstream) stream, (CUdeviceptr) (ctx->d_mbox + MBOX_FLAG), seq, flags);
    if (r != CUDA_SUCCESS) return fail(ctx, GPSIQ_ERR_CUDA, "gpsiq_mailbox_recv: cuStreamWaitValue64 failed", cudaSuccess);
    trace_mark(ctx, (cudaStream_t) stream, "mbox_recv(flag seen)");
    CU(cudaMemcpyAsync(ctx->d_carr_state, ctx->d_mbox + (seq & 1) * MBOX_SLOT, ctx->C * sizeof(double),
                       cudaMemcpyDeviceToDevice, (cudaStream_t) stream));
    trace_mark(ctx, (cudaStream_t) stream, "mbox_recv(copied)");
    return GPSIQ_OK;
}

int gpsiq_get_carrier(gpsiq_ctx* ctx, double* p) {
    if (!ctx || !p) return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_get_carrier: bad argument", cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(p, ctx->d_carr_state, ctx->C * sizeof(double), cudaMemcpyDeviceToHost));
    return GPSIQ_OK;
}

int gpsiq_set_carrier(gpsiq_ctx* ctx, const double* p) {
    if (!ctx || !p) return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_set_carrier: bad argument", cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(ctx->d_carr_state, p, ctx->C * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_est_state, p, ctx->C * sizeof(double), cudaMemcpyHostToDevice));
    return GPSIQ_OK;
}

int gpsiq_get_carrier_trace(gpsiq_ctx* ctx, double* trace, int n_epochs) {
    if (!ctx || !trace || n_epochs < 1 || n_epochs > ctx->E)
        return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_get_carrier_trace: bad argument", cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(trace, ctx->d_carr_trace, (size_t) n_epochs * ctx->C * sizeof(double), cudaMemcpyDeviceToHost));
    return GPSIQ_OK;
}

int16_t* gpsiq_device_iq(gpsiq_ctx* ctx) { return ctx ? ctx->d_iq : NULL; }

int gpsiq_checksum_device(gpsiq_ctx* ctx, const int16_t* iq_dev, int n_epochs, uint64_t* sums_out) {
    if (!ctx || !iq_dev || !sums_out || n_epochs < 1 || n_epochs > ctx->E)
        return fail(ctx, GPSIQ_ERR_ARG, "gpsiq_checksum_device: bad argument", cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemsetAsync(ctx->d_sums, 0, (size_t) n_epochs * 8, ctx->stream));
    dim3 grid(64, n_epochs);
    k_checksum<<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<const uint32_t*>(iq_dev), ctx->d_sums, ctx->N);
    ctx->launches += 1;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(sums_out, ctx->d_sums, (size_t) n_epochs * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return GPSIQ_OK;
}

int64_t gpsiq_launch_count(const gpsiq_ctx* ctx) { return ctx ? ctx->launches : 0; }

int gpsiq_carrier_fallbacks(gpsiq_ctx* ctx, int64_t* count) {
    if (!ctx || !count) return GPSIQ_ERR_ARG;
    int h = 0;
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(&h, ctx->d_fallbacks, sizeof h, cudaMemcpyDeviceToHost));
    *count = h;
    return GPSIQ_OK;
}

int gpsiq_slice_stats(gpsiq_ctx* ctx, int64_t* translated, int64_t* serial) {
    if (!ctx || !translated || !serial) return GPSIQ_ERR_ARG;
    unsigned long long h[2] = {0, 0};
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(h, ctx->d_slice_stats, sizeof h, cudaMemcpyDeviceToHost));
    *translated = (int64_t) h[0];
    *serial = (int64_t) h[1];
    return GPSIQ_OK;
}

int gpsiq_device_status(gpsiq_ctx* ctx) {
    if (!ctx) return GPSIQ_ERR_ARG;
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaDeviceSynchronize());
    return check_device_error(ctx);
}

int gpsiq_timing_begin(gpsiq_ctx* ctx) {
    if (!ctx) return GPSIQ_ERR_ARG;
    ctx->ev_count = 0;
    return GPSIQ_OK;
}

int gpsiq_timing_collect(gpsiq_ctx* ctx, int* n_steps, float* scan_ms, float* synth_ms) {
    if (!ctx) return GPSIQ_ERR_ARG;
    CU(cudaSetDevice(ctx->cfg.device));
    float a = 0.f, b = 0.f;
    for (int i = 0; i < ctx->ev_count; i++) {
        float t = 0.f;
        CU(cudaEventSynchronize(ctx->ev[i][2]));
        // (the scan phase of a batch submitted ahead runs on another stream and is not bracketed)
        if (cudaEventElapsedTime(&t, ctx->ev[i][0], ctx->ev[i][1]) == cudaSuccess && t > 0.f) a += t;
        else cudaGetLastError();
        CU(cudaEventElapsedTime(&t, ctx->ev[i][1], ctx->ev[i][2]));
        b += t;
    }
    if (n_steps) *n_steps = ctx->ev_count;
    if (scan_ms) *scan_ms = a;
    if (synth_ms) *synth_ms = b;
    return GPSIQ_OK;
}

// Re-launch the last k_synth_line (same anchors, same output range -- it rewrites identical samples)
// `reps` times back to back on an otherwise idle device and return the mean duration: the kernel ALONE.
int gpsiq_timing_sample_kernel_isolated(gpsiq_ctx* ctx, int reps, float* kernel_ms, int* epochs_per_launch) {
    if (!ctx || reps < 1 || !kernel_ms) return GPSIQ_ERR_ARG;
    if (!ctx->use_line || !ctx->last_ln.desc) return fail(ctx, GPSIQ_ERR_ARG, "no k_synth_line launch to repeat", cudaSuccess);
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaDeviceSynchronize());
    use_set(ctx, ctx->last_ln.set);
    const int C = ctx->C, N = ctx->N, ntiles = ctx->ntiles;
    cudaEvent_t e0 = ctx->ev[TIMING_RING - 1][3], e1 = ctx->ev[TIMING_RING - 1][4];
    for (int i = 0; i < reps + 1; i++) {  // first launch is a warm-up
        if (i == 1) CU(cudaEventRecord(e0, ctx->stream));
        const int ne = ctx->last_ln.ne, le0 = ctx->last_ln.e0;
        int grid = ne * ((ntiles + LN_UNIT - 1) / LN_UNIT);
        if (ctx->line_grid_cap > 0 && grid > ctx->line_grid_cap) grid = ctx->line_grid_cap;
        k_synth_line<<<grid, LN_THREADS, ln_smem_bytes(C), ctx->stream>>>(
            ctx->last_ln.desc, ctx->d_lutp + (size_t) le0 * C * 512, ctx->d_chips4,
            ctx->d_anch[ctx->last_ln.set] + (size_t) le0 * ntiles * C, ctx->d_flags + le0, ctx->d_flags + ctx->E + le0,
            ctx->last_ln.iq, ne, C, N, ntiles, ctx->cfg.carrier_mode == GPSIQ_CARRIER_INT32, ctx->d_err);
    }
    CU(cudaEventRecord(e1, ctx->stream));
    // the re-launches rewrote the samples WITHOUT the batch's patches (k_line_apply): put them back, so that the
    // output buffer still holds the exact stream afterwards (the patch list of the last rendered batch is intact)
    k_line_apply<<<4, 128, 0, ctx->stream>>>(ctx->d_patches, ctx->d_line_counters, ctx->patch_cap,
                                              reinterpret_cast<uint32_t*>(ctx->last_ln.iq) - (size_t) ctx->last_ln.e0 * N,
                                              (unsigned long long) ctx->last_ln.e0 * N,
                                              (unsigned long long) (ctx->last_ln.e0 + ctx->last_ln.ne) * N, NULL);
    CU(cudaEventSynchronize(e1));
    CU(cudaStreamSynchronize(ctx->stream));
    float t = 0.f;
    CU(cudaEventElapsedTime(&t, e0, e1));
    ctx->launches += reps + 2;
    *kernel_ms = t / reps;
    if (epochs_per_launch) *epochs_per_launch = ctx->last_ln.ne;
    return GPSIQ_OK;
}

int gpsiq_timing_sample_kernel(gpsiq_ctx* ctx, int* n_launches, float* kernel_ms, int* epochs_per_launch) {
    if (!ctx) return GPSIQ_ERR_ARG;
    CU(cudaSetDevice(ctx->cfg.device));
    float a = 0.f;
    int n = 0;
    if (ctx->use_line)
        for (int i = 0; i < ctx->ev_count; i++) {
            float t = 0.f;
            CU(cudaEventSynchronize(ctx->ev[i][4]));
            CU(cudaEventElapsedTime(&t, ctx->ev[i][3], ctx->ev[i][4]));
            a += t;
            n++;
        }
    if (n_launches) *n_launches = n;
    if (kernel_ms) *kernel_ms = a;
    if (epochs_per_launch) *epochs_per_launch = ctx->fixed_epochs;
    return GPSIQ_OK;
}

int gpsiq_trace_dump(gpsiq_ctx* ctx, int reset) {
    if (!ctx) return GPSIQ_ERR_ARG;
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaDeviceSynchronize());
    for (int i = 0; i < ctx->trace_n; i++) {
        float t = 0.f;
        cudaEventElapsedTime(&t, ctx->trace[0].ev, ctx->trace[i].ev);
        fprintf(stderr, "trace dev %d %9.3f ms  stream %d  %s  (enqueued by the host at %.3f ms)\n", ctx->cfg.device, t,
                ctx->trace[i].stream_id, ctx->trace[i].label, ctx->trace[i].host_ms - ctx->trace[0].host_ms);
    }
    if (reset) ctx->trace_n = 0;
    return GPSIQ_OK;
}

int gpsiq_line_stats(gpsiq_ctx* ctx, int64_t* hazard_tiles, int64_t* patches, int64_t* flagged_chunks) {
    if (!ctx) return GPSIQ_ERR_ARG;
    unsigned long long h[4] = {0, 0, 0, 0};
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaDeviceSynchronize());
    if (ctx->use_line) CU(cudaMemcpy(h, ctx->d_line_totals, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (hazard_tiles) *hazard_tiles = (int64_t) h[0];
    if (patches) *patches = (int64_t) h[1];
    if (flagged_chunks) *flagged_chunks = (int64_t) h[2];
    return GPSIQ_OK;
}

uint64_t gpsiq_minmod_host(uint64_t b, uint64_t a, uint64_t m, uint64_t n, uint64_t stop) {
    return minmod(b, a, m, n, stop);
}

int gpsiq_line_probe_host(int mode, double x0, double step, int n, int64_t* max_dev, int* mismatches, int* hazard) {
    if ((mode != NCO_CODE && mode != NCO_CARRIER) || n < 1 || n > (1 << 20)) return GPSIQ_ERR_ARG;
    const int B = (mode == NCO_CARRIER) ? LN_FBITS : LN_GBITS;
    const uint64_t A = (mode == NCO_CARRIER) ? ln_carr_fixed(x0) : ln_code_fixed(x0);
    const uint64_t d = (mode == NCO_CARRIER) ? ln_carr_slope(step) : ln_code_slope(step);
    double x = x0;
    uint64_t L = A;
    int wraps = 0, mism = 0;
    int64_t dev = 0;
    for (int i = 0; i < n; i++) {
        // truth in the line's coordinates: carrier mod 2^64, code unwrapped by 1023 chips per wrap
        const uint64_t T = (mode == NCO_CARRIER) ? ln_carr_fixed(x)
                                                 : ln_code_fixed(x) + (((uint64_t) wraps * 1023u) << LN_GBITS);
        int64_t dv = (int64_t) (T - L);
        if (mode == NCO_CARRIER && x == 1.0) dv = (int64_t) (0 - L);  // 1.0 saturates to 2^64 - 1: compare with 2^64
        if (dv < 0) dv = -dv;
        if (dv > dev) dev = dv;
        // the index k_synth_line evaluates for this sample (split-word line from the lane's exact start)
        uint32_t ci, gi;
        ln_kernel_index(A, A, d, d, (uint32_t) i, ci, gi);
        const uint32_t ki = (mode == NCO_CARRIER) ? ci : gi;
        const uint32_t ti = (mode == NCO_CARRIER) ? (uint32_t) (T >> B) & 511u : (uint32_t) (T >> B);
        if (ti != ki) mism++;
        L += d;
        if (mode == NCO_CARRIER) nco_step<NCO_CARRIER>(x, step, wraps);
        else nco_step<NCO_CODE>(x, step, wraps);
    }
    // the window k_line_anchor uses: truth within +-eps of the line, the kernel's value up to LN_K* below it
    const int64_t eps = ln_eps(mode == NCO_CARRIER, n);
    const int64_t K = (mode == NCO_CARRIER) ? LN_KF : LN_KG;
    if (max_dev) *max_dev = dev;
    if (mismatches) *mismatches = mism;
    if (hazard) *hazard = line_hazard(A, d, B, (uint64_t) n, -eps - K, eps) ? 1 : 0;
    return GPSIQ_OK;
}

// Host run of the line kernel's index arithmetic over whole epochs of real descriptors (tests): for every
// (epoch, slot, 1024-sample tile) the exact tile-start state comes from the literal recurrences
// (plutogpssim.c:2709-2713, 2741-2746), the tile is checked exactly as k_line_anchor checks it at tile level, and
// every sample's carrier-table / chip index as k_synth_line evaluates it (ln_kernel_index) is compared with the
// recurrence's.  bad counts differing samples in tiles the check CLEARED: it must stay 0 (samples of flagged tiles
// are repaired by k_line_patch).  carr0: the slots' phases before the first epoch (NULL: zeros).
int gpsiq_line_verify_host(const gpsiq_chan_desc* desc, int n_epochs, int C, int N, const double* carr0, int64_t* tiles_out,
                           int64_t* flagged_out, int64_t* bad_out, int64_t* lag_flagged_out) {
    if (!desc || n_epochs < 0 || C < 1 || C > GPSIQ_MAX_CHAN || N < 1) return GPSIQ_ERR_ARG;
    int64_t tiles = 0, flagged = 0, bad = 0, lag = 0;
    for (int c = 0; c < C; c++) {
        double ph = carr0 ? carr0[c] : 0.0;
        for (int e = 0; e < n_epochs; e++) {
            const gpsiq_chan_desc& d = desc[(size_t) e * C + c];
            if (d.prn <= 0) continue;
            if (d.flags & GPSIQ_FLAG_RESET_CARRIER) ph = d.carr_phase0;
            if (!(d.code_step > 0.0 && d.code_step <= 0.5) || !(fabs(d.carr_step) <= 0x1p-8)) return GPSIQ_ERR_ARG;  // line kernel's contract
            double cp = d.code_phase0;
            const uint64_t dF = ln_carr_slope(d.carr_step), dG = ln_code_slope(d.code_step);
            const uint64_t G0 = ln_code_fixed(d.code_phase0);
            int wraps = 0;  // code-period wraps of the recurrence since the epoch's first sample
            for (int t0 = 0; t0 < N; t0 += LN_TILE) {
                const int len = (N - t0 < LN_TILE) ? N - t0 : LN_TILE;
                // carrier anchor: the exact tile-start phase; code anchor: closed form on the epoch's line (k_line_anchor)
                const uint64_t FA = ln_carr_fixed(ph);
                uint64_t GA;
                int wl;
                ln_code_line(G0, dG, (uint32_t) t0, GA, wl);
                const int64_t eF = ln_eps(1, LN_TILE), eG = ln_eps(0, (int64_t) t0 + len);
                const bool hz = line_hazard(FA, dF, LN_FBITS, (uint64_t) len, -eF - LN_KF, eF) ||
                                line_hazard(GA, dG, LN_GBITS, (uint64_t) len, -eG - LN_KG, eG);
                int differ = 0;
                for (int n = 0; n < len; n++) {
                    int it = (int) floor(ph * 512.0);
                    if (it > 511) it = 511;
                    // chips since the code period the LINE is in at the tile start (G is not wrapped inside a tile)
                    const uint32_t chip_true = (uint32_t) ((int) cp + 1023 * (wraps - wl));
                    uint32_t ci, gi;
                    ln_kernel_index(FA, GA, dF, dG, (uint32_t) n, ci, gi);
                    if (ci != (uint32_t) it || gi != chip_true) differ++;
                    nco_step<NCO_CODE>(cp, d.code_step, wraps);
                    int w2 = 0;
                    nco_step<NCO_CARRIER>(ph, d.carr_step, w2);
                }
                tiles++;
                if (hz) { flagged++; lag += differ; }
                else bad += differ;
            }
        }
    }
    if (tiles_out) *tiles_out = tiles;
    if (flagged_out) *flagged_out = flagged;
    if (bad_out) *bad_out = bad;
    if (lag_flagged_out) *lag_flagged_out = lag;
    return GPSIQ_OK;
}

}  // extern "C"
