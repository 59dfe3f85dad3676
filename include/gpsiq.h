/* gpsiq.h — C-ABI of the B200 GPS L1 C/A baseband I/Q synthesizer.
 *
 * The reference (Mictronics/pluto-gps-sim) has no plugin / operator / FFI
 * interface for its hot path: the per-sample loop is inline in main()
 * between pthread_mutex_lock (plutogpssim.c:2689) and pthread_cond_signal
 * (plutogpssim.c:2757).  The seam is therefore defined by DATA:
 *
 *   inputs   the chan[MAX_CHAN] fields the loop reads {prn, ca, f_carr, f_code,
 *            carr_phase, code_phase, dwrd, iword, ibit, icode, dataBit, codeCA},
 *            gain[MAX_CHAN] and delt — all written by the per-epoch refresh at
 *            plutogpssim.c:2656-2687 (computeCodePhase, plutogpssim.c:1754-1787)
 *            and, for carr_phase, by allocateChannel (plutogpssim.c:1964);
 *   outputs  iq_buff[2*NUM_SAMPLES] (plutogpssim.c:2754-2755) and the advanced
 *            chan[i].carr_phase (plutogpssim.c:2741-2746), the only loop state
 *            the reference reads again in the next epoch.
 *
 * A drop-in replaces plutogpssim.c:2690-2756 by gpsiq_make_desc() per active
 * channel + gpsiq_synth() into iq_buff (INTEGRATION.md shows the patch).
 *
 * Conventions: plain C, int status returns (0 = GPSIQ_OK, negative = error,
 * text via gpsiq_strerror / gpsiq_last_error), caller-owned buffers, one
 * context per CUDA device, calls on one context are not re-entrant, different
 * contexts may be driven from different threads.  There is NO CPU fallback:
 * without a usable CUDA device gpsiq_create fails with GPSIQ_ERR_CUDA.
 */
#ifndef GPSIQ_H
#define GPSIQ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPSIQ_OK 0
#define GPSIQ_ERR_ARG (-1)      /* bad argument / out-of-contract descriptor */
#define GPSIQ_ERR_CUDA (-2)     /* CUDA runtime error (see gpsiq_last_error) */
#define GPSIQ_ERR_NOMEM (-3)
#define GPSIQ_ERR_CAPACITY (-4) /* n_epochs > cfg.max_epochs */

/* carrier NCO arithmetic */
#define GPSIQ_CARRIER_FLOAT 0 /* shipped build: FLOAT_CARR_PHASE defined, plutogpssim.h:12; binary64 recurrence plutogpssim.c:2697, 2741-2746 */
#define GPSIQ_CARRIER_INT32 1 /* compiled-out #else branch: uint32 phase, plutogpssim.c:2699, 2748, 1966-1967, 2675 */

/* gpsiq_chan_desc.flags */
#define GPSIQ_FLAG_RESET_CARRIER 1u /* slot was (re)allocated: start from carr_phase0 (plutogpssim.c:1964) */

/* synthesis kernel selection (gpsiq_config.kernel) */
#define GPSIQ_KERNEL_AUTO 0
#define GPSIQ_KERNEL_LANE_PER_CHANNEL 1 /* warp = sample tile, lane = channel, true FP64 steps */
/* (2 was the segment-list kernel of the first design; retired -- gpsiq_create rejects it) */
#define GPSIQ_KERNEL_LINE 3             /* one fixed-point line per 1024-sample tile + exact safety check (AUTO picks it) */

#define GPSIQ_MAX_CHAN 32
#define GPSIQ_CA_LEN 1023

/* One per channel slot per epoch: the state the reference's loop starts from.
 * All doubles are bit copies of what the host computed — the kernels never
 * re-derive them.  64 bytes. */
typedef struct gpsiq_chan_desc {
    int32_t prn;         /* 1..32; <=0 = inactive slot                       plutogpssim.h:153 */
    int32_t ms0;         /* iword*600 + ibit*20 + icode at sample 0          plutogpssim.c:1769-1778 */
    uint64_t navbits;    /* bit k = NAV data bit number (ms0/20 + k), i.e.
                            (dwrd[b/30] >> (29 - b%30)) & 1, b = ms0/20 + k   plutogpssim.c:1781, 2732 */
    double code_phase0;  /* chips, [0,1023)                                  plutogpssim.c:1770 */
    double code_step;    /* f_code*delt evaluated in binary64 on the host    plutogpssim.c:2709 */
    double carr_step;    /* FLOAT: f_carr*delt (cycles/sample)               plutogpssim.c:2741
                            INT32: (int)round(512*65536*f_carr*delt)         plutogpssim.c:2675 */
    double carr_phase0;  /* used iff flags & GPSIQ_FLAG_RESET_CARRIER.
                            FLOAT: cycles in [0,1); INT32: the uint32 phase  plutogpssim.c:1964-1967 */
    double gain;         /* path_loss*ant_gain                               plutogpssim.c:2685 */
    uint32_t flags;
    uint32_t reserved;
} gpsiq_chan_desc;

typedef struct gpsiq_config {
    int32_t device;            /* CUDA device ordinal */
    int32_t max_chan;          /* slots per epoch: 12 (plutogpssim.h:21) .. GPSIQ_MAX_CHAN */
    int32_t samples_per_epoch; /* 300000 = NUM_SAMPLES, plutogpssim.c:43-44 */
    int32_t carrier_mode;      /* GPSIQ_CARRIER_* */
    int32_t max_epochs;        /* capacity of one gpsiq_synth call */
    int32_t tile_samples;      /* 0 = default; checkpoint / CTA tile length */
    int32_t kernel;            /* GPSIQ_KERNEL_* */
    int32_t reserved[9];       /* reserved[0]: 1 = serial carrier scan (cross-check of the parallel one);
                                  reserved[1]: test hooks of the line kernel (1 force chunk re-check, 2 force every
                                  tile through the literal-recurrence patch path, 4 perturb the anchors); rest 0 */
} gpsiq_config;

typedef struct gpsiq_ctx gpsiq_ctx;

int gpsiq_create(gpsiq_ctx **ctx, const gpsiq_config *cfg);
void gpsiq_destroy(gpsiq_ctx *ctx);

/* Synthesize n_epochs consecutive epochs.  desc is [n_epochs][max_chan] in
 * HOST memory; iq_out receives n_epochs*samples_per_epoch interleaved
 * little-endian int16 I,Q pairs in HOST memory (pinned memory from
 * gpsiq_host_alloc makes the copies asynchronous), or may be NULL to leave the
 * result in the context's device buffer (gpsiq_device_iq).  Blocking: returns
 * when iq_out is complete.  The carrier phase of every slot continues from the
 * previous call (or from gpsiq_set_carrier / a RESET_CARRIER descriptor).
 * Replaces plutogpssim.c:2690-2756 for n_epochs epochs at once. */
int gpsiq_synth(gpsiq_ctx *ctx, const gpsiq_chan_desc *desc, int n_epochs, int16_t *iq_out);

/* Same, all-device and asynchronous: desc_dev and iq_dev are DEVICE pointers
 * (iq_dev 16-byte aligned), work is enqueued on cuda_stream (a cudaStream_t, used
 * as given: NULL is the CUDA default stream) and the call returns without waiting. */
int gpsiq_synth_device(gpsiq_ctx *ctx, const gpsiq_chan_desc *desc_dev, int n_epochs, int16_t *iq_dev,
                       void *cuda_stream);

/* Streaming pair (the submit/fetch seam of SURVEY.md §8b): gpsiq_submit_device copies the
 * descriptors (DEVICE pointer; ordered after the work already enqueued on after_stream) and
 * runs every scan phase for the batch ahead of time on the context's own stream;
 * gpsiq_fetch_device renders the oldest submitted batch into iq_dev on cuda_stream.  Up to two
 * batches may be in flight, so the serial carrier chain of batch k+1 overlaps the sample
 * kernels of batch k.  Results are identical to gpsiq_synth_device called batch by batch.
 * Do not interleave with the other synthesis calls while batches are in flight. */
int gpsiq_submit_device(gpsiq_ctx *ctx, const gpsiq_chan_desc *desc_dev, int n_epochs, void *after_stream);
/* Host-buffer versions: gpsiq_submit copies the descriptors (the caller may reuse desc at once) and
 * returns immediately; gpsiq_fetch blocks until the oldest submitted batch is complete in iq_out
 * (pinned memory from gpsiq_host_alloc keeps the copies asynchronous).  The producer loop of the
 * reference's main (plutogpssim.c:2655-2806) maps onto: submit(k+1); fetch(k); push(k). */
int gpsiq_submit(gpsiq_ctx *ctx, const gpsiq_chan_desc *desc, int n_epochs);
int gpsiq_fetch(gpsiq_ctx *ctx, int16_t *iq_out);
int gpsiq_fetch_device(gpsiq_ctx *ctx, int16_t *iq_dev, void *cuda_stream);

/* One stream over several GPUs of THIS process (the reference's epoch loop, plutogpssim.c:2655-2806, over N devices;
 * time slices, the carrier phase handed from slice to slice).  Consecutive batches go to consecutive
 * devices cfg->device, cfg->device + 1, ...; the exact carrier chain of a batch starts from the state the previous
 * batch left on its device (two small asynchronous copies ordered by events; no host synchronisation, no kernel).
 * Same contract as gpsiq_submit / gpsiq_fetch: host descriptors in, host I/Q out, results identical to one context
 * fed batch by batch.  Up to 3 batches per device may be submitted ahead; gpsiq_multi_fetch_begin starts the rendering
 * and the device-to-host copies of the oldest submitted batch without waiting (at most two begun per device),
 * gpsiq_multi_fetch_end waits for the oldest begun one; gpsiq_multi_fetch = begin + end. */
typedef struct gpsiq_multi gpsiq_multi;
int gpsiq_multi_create(gpsiq_multi **m, const gpsiq_config *cfg, int n_devices);
void gpsiq_multi_destroy(gpsiq_multi *m);
int gpsiq_multi_submit(gpsiq_multi *m, const gpsiq_chan_desc *desc, int n_epochs);
int gpsiq_multi_fetch_begin(gpsiq_multi *m, int16_t *iq_out);
int gpsiq_multi_fetch_end(gpsiq_multi *m);
int gpsiq_multi_fetch(gpsiq_multi *m, int16_t *iq_out);
int gpsiq_multi_devices(const gpsiq_multi *m);
int64_t gpsiq_multi_launch_count(const gpsiq_multi *m);
const char *gpsiq_multi_last_error(const gpsiq_multi *m);

/* Carrier phase per slot after the last synthesized epoch (max_chan values;
 * INT32 mode: the uint32 phase as a double) — chan[i].carr_phase after
 * plutogpssim.c:2741-2748.  Used for time-slice hand-off between GPUs. */
int gpsiq_get_carrier(gpsiq_ctx *ctx, double *carr_phase);
int gpsiq_set_carrier(gpsiq_ctx *ctx, const double *carr_phase);
/* Post-epoch carrier phase of every (epoch, slot) of the last call
 * ([n_epochs][max_chan]); parity aid. */
int gpsiq_get_carrier_trace(gpsiq_ctx *ctx, double *trace, int n_epochs);

/* Device buffer holding the last gpsiq_synth result when iq_out was NULL. */
int16_t *gpsiq_device_iq(gpsiq_ctx *ctx);

/* Order-independent 64-bit checksum of each epoch of an I/Q stream in DEVICE
 * memory (sum over samples of mix64(index, I|Q<<16)); lets multi-GB runs be
 * compared with the oracle without a device-to-host copy of the stream.
 * sums_out: n_epochs values in HOST memory. */
int gpsiq_checksum_device(gpsiq_ctx *ctx, const int16_t *iq_dev, int n_epochs, uint64_t *sums_out);

/* Host helper: fold the reference's per-channel loop inputs into a descriptor.
 * Arguments are the chan[i] fields after computeCodePhase (plutogpssim.c:1754-1787),
 * gain[i] (plutogpssim.c:2685) and delt (plutogpssim.c:2397).  dwrd is the
 * reference's 60-word NAV buffer (30-bit words, plutogpssim.h:166) passed as
 * 64-bit values ('unsigned long' on LP64).  carr_phase_is_new != 0 marks a slot
 * whose carr_phase was just initialised by allocateChannel.  Pure host code. */
int gpsiq_make_desc(gpsiq_chan_desc *out, int carrier_mode, int prn, double f_carr, double f_code, double delt,
                    double carr_phase, double code_phase, const uint64_t *dwrd60, int iword, int ibit, int icode,
                    double gain, int carr_phase_is_new);

/* Host execution of the exact NCO fast-forward the scan kernels run on the
 * device (csrc/nco_scan.cuh): advances *phase by `count` steps of the code
 * (mode 0: plutogpssim.c:2709-2713) or carrier (mode 1: plutogpssim.c:2741-2746)
 * recurrence and adds the number of code wraps to *wraps.  Diagnostic / test aid. */
int gpsiq_nco_advance(int mode, double *phase, double step, int64_t count, int64_t *wraps);

/* Host execution of the parallel carrier scan (speculate / translate / verify,
 * csrc/nco_scan.cuh) for one channel over n_epochs epochs with per-epoch steps:
 * writes the exact phase at every tile start ([n_epochs][ceil(N/T)]), the final
 * phase and the number of epochs that fell back to the serial scan.  est_err is
 * added to the start-phase estimate of every epoch (0 = what the device does) so
 * tests can force both the translated and the fallback path.  Diagnostic aid. */
int gpsiq_carrier_chain_host(const double *steps, int n_epochs, int N, int T, double x0, double est_err,
                             double *ck_out, double *x_end_out, int *n_fallback);

/* The same through the fifth level (csrc/nco_scan.cuh: slice_chain_group / slice_verify), as the device chains a
 * batch: the groups chained speculatively from the ESTIMATED batch start, ONE exact head scan matched against that
 * trajectory, then every group chained on its own from its translated start phase (the device does these in parallel,
 * after the hand-off of a time-sliced run).  *how_out: 1 = passed by translation, 0 = chained serially (no usable
 * slice-level speculation), -2 = a group ended somewhere else than the translation predicts (internal error).
 * ties_out (2 ints, may be NULL): [0] usable group trajectories that contain a tie-wrap of a tie-capable step
 * (csrc/nco_scan.cuh: TieEvent), [1] translations that changed at such an event (groups + 1000 * slice level).
 * Diagnostic aid.
 * flags / phase0 (may be NULL): per epoch, bit 0 = the slot is inactive in that epoch (the phase passes through),
 * bit 1 = the slot is re-seeded with phase0[e] at its first sample (a slice with a re-seed is chained serially). */
int gpsiq_carrier_slice_host(const double *steps, int n_epochs, int N, int T, double x0, double est_err,
                             double *ck_out, double *x_end_out, int *n_fallback, int *how_out, int *ties_out,
                             const int *flags, const double *phase0);
/* Study aid (tools/margin_study.py): the same run with the device's group size (group_epochs, 64 on the device) and a
 * residual rate subtracted from every closed-form epoch advance (what k_prepare does with its measured rate), reporting
 * out5 = { decision margin of the slice-level trajectory, variant 0; variant 1; closed-form end estimate minus the exact
 * end phase (cycles); smallest usable group-level margin; number of unusable groups }.  Diagnostic aid. */
int gpsiq_carrier_study_host(const double *steps, int n_epochs, int N, int T, double x0, double est_err, double est_rate,
                             int group_epochs, double *out5, int *how_out, int *n_fallback);

/* Pinned host memory for descriptors / I/Q (cudaHostAlloc). */
void *gpsiq_host_alloc(size_t bytes);
void gpsiq_host_free(void *p);

/* The phases of gpsiq_synth_device, separately, for time-sliced multi-GPU runs
 * (pluto_gps_sim_b200/timeslice.py).  A batch goes through
 *   prepare   -> LUTs, binade tables, and the batch's closed-form phase advance
 *                per slot (advance_dev, 2*max_chan doubles: advance or absolute
 *                phase, then 0/1 "re-seeded" flags), which the owners of LATER
 *                slices fold into their start-phase estimates;
 *   speculate -> code-NCO scan + speculative carrier scans from the context's
 *                start-phase ESTIMATE; needs no exact phase, so all ranks run it
 *                at the same time;
 *   chain     -> the only serial part: the exact carrier phase is chained through
 *                the batch from the context's carrier state (received from the
 *                previous slice's owner); O(one carrier cycle) per epoch;
 *   render    -> the per-sample synthesis.
 * gpsiq_scan_device = prepare + speculate + chain.  Estimates only affect speed:
 * a bad one makes epochs fall back to the serial scan, never changes a result.
 * Batches move through the phases in order, each phase with its own cursor over the context's three scan sets:
 * prepare / speculate of later batches may be enqueued (on another stream) before the chain of an earlier one --
 * time-sliced runs speculate slice k+1 while the ring still carries the exact phases of slice k.  With the chain on
 * another stream than the speculation, set GPSIQ_OPT_CHAIN_KEEPS_ESTIMATE. */
int gpsiq_prepare_device(gpsiq_ctx *ctx, const gpsiq_chan_desc *desc_dev, int n_epochs, double *advance_dev,
                         void *cuda_stream);
int gpsiq_speculate_device(gpsiq_ctx *ctx, const gpsiq_chan_desc *desc_dev, int n_epochs, void *cuda_stream);
int gpsiq_chain_device(gpsiq_ctx *ctx, const gpsiq_chan_desc *desc_dev, int n_epochs, void *cuda_stream);
int gpsiq_scan_device(gpsiq_ctx *ctx, const gpsiq_chan_desc *desc_dev, int n_epochs, void *cuda_stream);
int gpsiq_render_device(gpsiq_ctx *ctx, const gpsiq_chan_desc *desc_dev, int n_epochs, int16_t *iq_dev,
                        void *cuda_stream);
/* estimate <- fold(estimate, advance): skip over a slice synthesized elsewhere */
int gpsiq_estimate_fold_device(gpsiq_ctx *ctx, const double *advance_dev, void *cuda_stream);
/* Integer carrier only: carrier state <- fold(state, advance), EXACT (uint32 arithmetic, plutogpssim.c:2748): skips
 * over a slice synthesized elsewhere, where advance is what gpsiq_prepare_device returned for that slice (phase
 * advance modulo 2^32, or an absolute phase when the slice re-seeded the slot).  With it the time slices of the
 * integer-carrier mode need no ring: every rank folds the advances of the slices before its own. */
int gpsiq_carrier_fold_device(gpsiq_ctx *ctx, const double *advance_dev, void *cuda_stream);
/* estimate <- max_chan doubles in DEVICE memory (e.g. an exact phase saved earlier with gpsiq_carrier_to_device, to be
 * advanced with gpsiq_estimate_fold_device over the slices since) */
int gpsiq_estimate_from_device(gpsiq_ctx *ctx, const double *src_dev, void *cuda_stream);
/* estimate <- the context's exact carrier state (e.g. right after gpsiq_carrier_from_device) */
int gpsiq_estimate_anchor_device(gpsiq_ctx *ctx, void *cuda_stream);
/* Copy the carrier state (max_chan doubles) to / from DEVICE memory on a stream
 * (e.g. the buffer of an NCCL send / recv). */
/* Time-sliced runs keep the start-phase ESTIMATE on its own (speculation) stream, decoupled from the exact
 * chain: with GPSIQ_OPT_CHAIN_KEEPS_ESTIMATE set, gpsiq_chain_device no longer re-anchors the estimate on the
 * exact phase; gpsiq_estimate_to_device saves it, and gpsiq_estimate_correct_device feeds back
 * gain * (an older saved estimate - the exact phase later found for the same instant).  Estimates only ever
 * affect speed (how often the serial fallback runs), never a sample. */
#define GPSIQ_OPT_CHAIN_KEEPS_ESTIMATE 1
/* Rendering a batch while another one has been scanned ahead normally waits only for that batch's chunk
 * speculation; with this option it waits for its exact chain as well (time-sliced multi-GPU runs: the chain is a
 * hop of the inter-GPU ring and must not compete with the sample kernel for the SMs). */
#define GPSIQ_OPT_RENDER_AFTER_NEXT_CHAIN 2
/* Most CTAs of one k_synth_line launch; the CTAs then stride over the launch's work units (persistent form).
 * 0 (default): one CTA per unit.  A cap just below 2 x the SM count leaves a few half-empty SMs to the small
 * latency-bound kernels that run beside the sample kernel (the next batch's chain, a ring hop). */
#define GPSIQ_OPT_LINE_GRID_CAP 3
/* Time-sliced runs whose speculation must not wait for the ring: the running start-phase estimate is no longer
 * re-anchored on the exact phase by the chain.  After every gpsiq_speculate_device it becomes the END of that
 * batch's slice-level speculative trajectory (an exact advance from the estimated start), the caller folds in the
 * advances of the slices other GPUs own (gpsiq_estimate_fold_device), and the accumulated error is corrected open
 * loop from what each exact chain measures later (csrc/gpsiq.cu: k_est_open_loop).  The chain of a batch may then run
 * on another stream, any time after its speculation. */
#define GPSIQ_OPT_FREE_RUNNING_ESTIMATE 4
int gpsiq_set_option(gpsiq_ctx *ctx, int option, int value);
int gpsiq_estimate_to_device(gpsiq_ctx *ctx, double *dst_dev, void *cuda_stream);
/* SM-free hand-off of the carrier state between the GPUs of one node (time-sliced runs; replaces the NCCL
 * send/recv of max_chan doubles at a slice boundary -- chan[i].carr_phase, plutogpssim.c:2741-2746, is the only
 * state that crosses it).  A mailbox is a small device buffer (two state slots + a 64-bit sequence flag) exported
 * as a 64-byte CUDA IPC handle.  The sender opens the NEXT rank's mailbox once; gpsiq_mailbox_send copies the
 * context's carrier state into slot (seq & 1) with the copy engine and then writes seq to the flag with a stream
 * memory operation; gpsiq_mailbox_recv makes the stream wait until the own flag >= seq and then loads slot
 * (seq & 1) into the carrier state.  No kernel runs on either GPU.  seq must increase by one per message on an
 * edge; a slot is reused two messages later (in a ring the receiver has consumed it by then). */
int gpsiq_mailbox_create(gpsiq_ctx *ctx, void *ipc_handle_out_64_bytes);
int gpsiq_mailbox_open(gpsiq_ctx *ctx, const void *ipc_handle_64_bytes, int peer_device);
int gpsiq_mailbox_send(gpsiq_ctx *ctx, uint64_t seq, void *cuda_stream);
int gpsiq_mailbox_recv(gpsiq_ctx *ctx, uint64_t seq, void *cuda_stream);
/* The hand-off FUSED into the chain kernel (one kernel per hop of a time-sliced run): the kernel waits for message
 * recv_seq in this context's mailbox (acquire load of the sequence flag the previous GPU writes over NVLink peer
 * memory), chains the batch from the phases it finds there, and stores its end phases as message send_seq into the
 * next GPU's mailbox (peer stores, system-scope fence, release store of the flag).  recv_seq == 0: start from the
 * context's carrier state (as gpsiq_chain_device); send_seq == 0: no send.  start_copy_dev (may be NULL) receives the
 * exact start phases.  The wait is bounded (~10 s): a message that never arrives sets the context's error word.
 * Replaces gpsiq_mailbox_recv + gpsiq_chain_device + gpsiq_mailbox_send; messages are interchangeable with theirs. */
int gpsiq_chain_handoff_device(gpsiq_ctx *ctx, const gpsiq_chan_desc *desc_dev, int n_epochs, uint64_t recv_seq,
                               uint64_t send_seq, double *start_copy_dev, void *cuda_stream);
int gpsiq_estimate_correct_device(gpsiq_ctx *ctx, const double *exact_old_dev, const double *est_old_dev, double gain,
                                  void *cuda_stream);
int gpsiq_carrier_to_device(gpsiq_ctx *ctx, double *dst_dev, void *cuda_stream);
int gpsiq_carrier_from_device(gpsiq_ctx *ctx, const double *src_dev, void *cuda_stream);

/* Number of CUDA kernel launches issued by this context so far. */
int64_t gpsiq_launch_count(const gpsiq_ctx *ctx);
/* (epoch, slot) pairs whose carrier scan fell back to the serial path so far. */
int gpsiq_carrier_fallbacks(gpsiq_ctx *ctx, int64_t *count);
/* Float carrier: (slot, batch) chains whose exact chain was ONE head scan + a translation of the slice-level
 * speculative trajectory (csrc/nco_scan.cuh, level 5), and the ones chained serially group by group (stream start,
 * re-seeded slots, a poor start-phase estimate).  Speed only: both give the same phases. */
int gpsiq_slice_stats(gpsiq_ctx *ctx, int64_t *translated, int64_t *serial);
/* Waits for the device and reports an error a kernel has flagged since the last check (an out-of-contract descriptor,
 * a table copy that never completed, an internal inconsistency of the carrier chain) -- what gpsiq_synth / gpsiq_fetch
 * check by themselves; for callers of the device-pointer API.  GPSIQ_OK if there is none. */
int gpsiq_device_status(gpsiq_ctx *ctx);
/* Device-side timing (CUDA events on the launching stream).  After
 * gpsiq_timing_begin every synth call records events around its scan phase and
 * its synthesis kernel (up to 64 calls); gpsiq_timing_collect waits for them and
 * returns the number of recorded calls and the summed durations in ms. */
int gpsiq_timing_begin(gpsiq_ctx *ctx);
int gpsiq_timing_collect(gpsiq_ctx *ctx, int *n_steps, float *scan_ms, float *synth_ms);
/* The dominant kernel alone: summed duration of the first k_synth_line launch of every
 * recorded call (CUDA events on the launching stream around that launch only), the
 * number of such launches and the epochs each one covers (epochs*samples_per_epoch*4
 * algorithmic bytes per launch). */
int gpsiq_timing_sample_kernel(gpsiq_ctx *ctx, int *n_launches, float *kernel_ms, int *epochs_per_launch);
/* The same kernel ALONE: re-launches the last k_synth_line (same inputs, same output range) `reps`
 * times back to back on an idle device and returns the mean duration.  In the pipelined calls the
 * kernel shares the SMs with the scan kernels of the next batch, so its in-pipeline duration says
 * little about the kernel itself. */
int gpsiq_timing_sample_kernel_isolated(gpsiq_ctx *ctx, int reps, float *kernel_ms, int *epochs_per_launch);

/* Line kernel diagnostics, accumulated over the context's life: (tile, slot) pairs that had to be
 * re-checked with the literal recurrence, samples patched, (epoch, slot) pairs the epoch-level check could not
 * clear (refined per 32-tile chunk, then per tile). */
int gpsiq_line_stats(gpsiq_ctx *ctx, int64_t *hazard_tiles, int64_t *patches, int64_t *flagged_chunks);
/* Host execution of the safety check's arithmetic (tests):
 * gpsiq_minmod_host: min over x in [0,n) of (b + a*x) mod m (may return any attained value < stop early).
 * gpsiq_line_probe_host: run the literal NCO recurrence (plutogpssim.c:2709-2713 / 2741-2746) for n samples
 * from x0 next to the straight fixed-point line of the line kernel; reports the largest deviation of the
 * recurrence from that line (line units), the number of samples whose table/chip index differs from the index
 * the kernel evaluates (its split-word form of the line, truncation included), and whether the check flags the run. */
uint64_t gpsiq_minmod_host(uint64_t b, uint64_t a, uint64_t m, uint64_t n, uint64_t stop);
int gpsiq_line_probe_host(int mode, double x0, double step, int n, int64_t *max_dev, int *mismatches, int *hazard);
/* The same over whole epochs of real descriptors: literal recurrences give the exact state at every 1024-sample
 * tile start, every tile is checked as the device checks it, every sample's table / chip index as the line kernel
 * evaluates it is compared with the recurrence's.  bad = differing samples in tiles the check cleared (must be 0);
 * lag_flagged = differing samples in flagged tiles (the patch path repairs those).  carr0: phases before the first
 * epoch (NULL: zeros). */
int gpsiq_line_verify_host(const gpsiq_chan_desc *desc, int n_epochs, int max_chan, int samples_per_epoch,
                           const double *carr0, int64_t *tiles, int64_t *flagged, int64_t *bad, int64_t *lag_flagged);

/* Diagnostics: with GPSIQ_TRACE set in the environment the context records a CUDA event after every kernel it
 * launches; this prints them to stderr as milliseconds since the first one (stream 0 = caller / render stream,
 * 1 = scan stream, 2 = code-scan side stream).  reset != 0 clears the record. */
int gpsiq_trace_dump(gpsiq_ctx *ctx, int reset);

const char *gpsiq_strerror(int status);
const char *gpsiq_last_error(const gpsiq_ctx *ctx);
const char *gpsiq_version(void);

/* The 512-entry carrier tables and the C/A code the kernels use (host copies),
 * for parity checks against plutogpssim.c:93-161 and plutogpssim.c:207-244. */
void gpsiq_get_tables(int32_t *sin512, int32_t *cos512);
int gpsiq_get_ca_code(int prn, uint8_t *chips1023);

#ifdef __cplusplus
}
#endif
#endif /* GPSIQ_H */
