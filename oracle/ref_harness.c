/* TEST INFRASTRUCTURE — single-thread harness around the UNMODIFIED reference.
 *
 * This wrapper compiles /root/reference/plutogpssim.c where it lies (by
 * #include, no copy, no textual patch) and turns its two-thread SDR program
 * into a deterministic batch oracle:
 *
 *   - the reference's main() is renamed ref_main and called from ours;
 *   - pthread_create/join become no-ops, so the libiio TX thread
 *     (plutogpssim.c:2058-2190) never runs;
 *   - pthread_mutex_lock(), which the reference calls immediately before its
 *     per-sample loop (plutogpssim.c:2689), becomes a hook that snapshots the
 *     per-epoch channel state the loop is about to consume (chan[], gain[],
 *     delt — all locals of the reference's main; captured by name through the
 *     macro, with file-scope decoys of the same names for the other call
 *     sites);
 *   - pthread_cond_wait(), which the reference calls right after the loop
 *     (plutogpssim.c:2758), becomes a hook that takes iq_buff as the finished
 *     epoch, records the post-loop carr_phase and the loop's wall time, and
 *     raises plutotx.exit after $REF_EPOCHS epochs.
 *
 * The sample loop itself, and every function feeding it, is the reference's
 * own code compiled from its own source.  Outputs (all optional, by env var):
 *   REF_IQ_OUT    raw interleaved little-endian int16 I,Q, 300 000 samples/epoch
 *   REF_DESC_OUT  one ref_dump_t per (epoch, slot)          [layout below]
 *   REF_EPOCHS    number of 0.1 s epochs to generate (default 10)
 *   REF_NO_PIN    (always on here) thread_to_core() pinning is neutralised so
 *                 that several replicas can be timed on different cores.
 * A one-line JSON summary goes to stdout.
 *
 * -DORACLE_MAX_CHAN=32 rebuilds the same source with 32 channel slots
 * (MAX_CHAN is only used for array extents and loop bounds in the reference's
 * main/allocateChannel; plutogpssim.h:21).
 *
 * Nothing under oracle/ is part of the shipped product path.
 */
#define _GNU_SOURCE
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include <pthread.h>
#include <errno.h>
#include <signal.h>
#include <sched.h>
#include <unistd.h>
#include <curl/curl.h>
#include <iio.h>
#include <ad9361.h>
#include <zlib.h>

#include "plutogpssim.h"

#ifdef ORACLE_MAX_CHAN
#undef MAX_CHAN
#define MAX_CHAN (ORACLE_MAX_CHAN)
#endif

/* Layout shared with tools/refdump.py (keep in sync). 384 bytes. */
typedef struct {
    int32_t epoch, slot, prn, iword, ibit, icode, dataBit, codeCA;
    double f_carr, f_code, delt, carr_phase, code_phase, gain;
    double carr_phase_end;
    double azel[2];
    double rho_range, rho_d, rho_iono;
    double g0_sec;
    int32_t g0_week, pad;
    uint32_t dwrd[N_DWRD];
} ref_dump_t;

static int oracle_loop_begin(const void *chan_v, const double *gain_v, double delt_v);
static int oracle_loop_end(const void *chan_v);

/* Decoys: the hooked calls also occur in the TX thread and the signal handler,
 * where the reference's locals of these names are not in scope. */
static const void *const chan = NULL;
static const double *const gain = NULL;
static const double delt = 0.0;

#define pthread_create(a, b, c, d) (0)
#define pthread_join(a, b) (0)
#define pthread_setaffinity_np(a, b, c) (0)
#define pthread_mutex_lock(m) oracle_loop_begin((const void *) chan, (const double *) gain, delt)
#define pthread_cond_wait(c, m) oracle_loop_end((const void *) chan)
#define pthread_cond_signal(c) (0)
#define main ref_main

#include "plutogpssim.c"

#undef main
#undef pthread_mutex_lock
#undef pthread_cond_wait
#undef pthread_cond_signal

static FILE *h_iq, *h_desc;
static long h_epochs = 10, h_done;
static ref_dump_t h_rec[MAX_CHAN];
static struct timespec h_t0;
static double h_loop_seconds;
static int h_nchan_max;

static int oracle_loop_begin(const void *chan_v, const double *gain_v, double delt_v) {
    const channel_t *ch = chan_v;
    if (!ch) return 0;                      /* TX-thread / handler call site */
    int nact = 0;
    for (int i = 0; i < MAX_CHAN; i++) {
        ref_dump_t *r = &h_rec[i];
        memset(r, 0, sizeof *r);
        r->epoch = (int32_t) h_done;
        r->slot = i;
        r->prn = ch[i].prn;
        if (ch[i].prn <= 0) continue;
        nact++;
        r->iword = ch[i].iword; r->ibit = ch[i].ibit; r->icode = ch[i].icode;
        r->dataBit = ch[i].dataBit; r->codeCA = ch[i].codeCA;
        r->f_carr = ch[i].f_carr; r->f_code = ch[i].f_code; r->delt = delt_v;
#ifdef FLOAT_CARR_PHASE
        r->carr_phase = ch[i].carr_phase;
#else
        r->carr_phase = (double) ch[i].carr_phase;
        r->carr_phase_end = (double) ch[i].carr_phasestep;   /* overwritten below; step re-derived by the reader */
#endif
        r->code_phase = ch[i].code_phase;
        r->gain = gain_v[i];
        r->azel[0] = ch[i].azel[0]; r->azel[1] = ch[i].azel[1];
        r->rho_range = ch[i].rho0.range; r->rho_d = ch[i].rho0.d; r->rho_iono = ch[i].rho0.iono_delay;
        r->g0_sec = ch[i].g0.sec; r->g0_week = ch[i].g0.week;
        for (int w = 0; w < N_DWRD; w++) r->dwrd[w] = (uint32_t) ch[i].dwrd[w];
    }
    if (nact > h_nchan_max) h_nchan_max = nact;
    clock_gettime(CLOCK_MONOTONIC, &h_t0);
    return 0;
}

static int oracle_loop_end(const void *chan_v) {
    const channel_t *ch = chan_v;
    if (!ch) return 0;
    struct timespec t1;
    clock_gettime(CLOCK_MONOTONIC, &t1);
    h_loop_seconds += (double) (t1.tv_sec - h_t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - h_t0.tv_nsec);
    for (int i = 0; i < MAX_CHAN; i++)
        if (h_rec[i].prn > 0) h_rec[i].carr_phase_end = (double) ch[i].carr_phase;
    if (h_iq) fwrite(iq_buff, 1, BUFFER_SIZE, h_iq);
    if (h_desc) fwrite(h_rec, sizeof(ref_dump_t), MAX_CHAN, h_desc);
    if (++h_done >= h_epochs) plutotx.exit = true;
    return 0;
}

int main(int argc, char *argv[]) {
    const char *s;
    if ((s = getenv("REF_EPOCHS"))) h_epochs = atol(s);
    if ((s = getenv("REF_IQ_OUT"))) h_iq = fopen(s, "wb");
    if ((s = getenv("REF_DESC_OUT"))) h_desc = fopen(s, "wb");
    if (getenv("REF_TABLES")) {            /* dump the carrier tables for fixture checks */
        for (int i = 0; i < 512; i++) printf("%d %d\n", sinTable512[i], cosTable512[i]);
        return 0;
    }
    int rc = ref_main(argc, argv);
    if (h_iq) fclose(h_iq);
    if (h_desc) fclose(h_desc);
    printf("{\"epochs\": %ld, \"samples_per_epoch\": %d, \"max_chan\": %d, \"active_chan_max\": %d, "
           "\"loop_seconds\": %.6f, \"msamples_per_s\": %.4f, \"dump_record_bytes\": %zu}\n",
           h_done, NUM_SAMPLES, MAX_CHAN, h_nchan_max, h_loop_seconds,
           h_loop_seconds > 0 ? 1e-6 * (double) h_done * NUM_SAMPLES / h_loop_seconds : 0.0, sizeof(ref_dump_t));
    return rc;
}
