/* gpssink.h -- C-ABI of the output transport (part of libgpshost.so, plain C++ / no CUDA).
 *
 * What happens to the I/Q stream AFTER the sample loop.  In the reference
 * that is one thread, pluto_tx_thread_ep (plutogpssim.c:2058-2190): it opens the ADALM-Pluto through
 * libiio, programs the AD9361, creates ONE non-cyclic buffer of NUM_SAMPLES = 300000 I/Q pairs backed by
 * 12 kernel buffers, and then forever copies the producer's buffer into it and pushes it
 * (plutogpssim.c:2146-2158).  Here the same contract sits behind a small interface:
 *
 *   push unit          300000 interleaved little-endian int16 I,Q pairs = 1 200 000 bytes  (plutogpssim.c:43-45)
 *   null sink          count and discard
 *   file sink          append to a file / stdout (what gps-sdr-sim, the reference's ancestor, offered)
 *   radio sink         the libiio sequence of the reference.  libiio / libad9361 are NOT link-time
 *                      dependencies: they are dlopen()ed when a radio sink is opened, so the library loads
 *                      on boxes without them and the open fails with GPSSINK_ERR_BACKEND instead.
 *
 * Every sink can be driven synchronously (gpssink_push) or through its own writer thread
 * (gpssink_submit / gpssink_wait): the producer hands over a whole batch of push units that lie in one
 * host buffer (the pinned buffer gpsiq_fetch filled) and goes on to fetch the next batch while the
 * writer drains this one -- the reference's mutex/condvar hand-off (plutogpssim.c:2147-2150, 2757-2758)
 * without its lost/duplicated buffers: every unit is pushed exactly once, in order. */
#ifndef GPSSINK_H
#define GPSSINK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPSSINK_OK 0
#define GPSSINK_ERR_ARG (-1)
#define GPSSINK_ERR_IO (-2)        /* file cannot be opened / short write */
#define GPSSINK_ERR_BACKEND (-3)   /* libiio (or a symbol of it) not available */
#define GPSSINK_ERR_DEVICE (-4)    /* no context / no Pluto TX device / buffer cannot be created */
#define GPSSINK_ERR_PUSH (-5)      /* the device refused a buffer (plutogpssim.c:2153-2156) */

#define GPSSINK_PUSH_PAIRS 300000  /* NUM_SAMPLES, plutogpssim.c:44 */

/* Radio parameters = the reference's plutotx block (plutogpssim.h, set at plutogpssim.c:2270-2276). */
typedef struct gpssink_radio_config {
    const char *uri;           /* -U  (plutogpssim.c:2377)  e.g. usb:1.2.5 */
    const char *hostname;      /* -N  (plutogpssim.c:2380)  default pluto.local when neither is given */
    double gain_db;            /* -A  TX attenuation, clamped to [-80, 0]; default -20 (plutogpssim.c:2274, 2367-2370) */
    long long bw_hz;           /* -B  RF bandwidth, clamped to [1, 5] MHz when given; default 2*TX_SAMPLE_FREQ (2270, 2372-2375) */
    long long fs_hz;           /* -s  baseband sample rate (plutogpssim.c:2271, 2325) */
    long long lo_hz;           /* 1575420000: L1 (plutogpssim.c:2272) */
    const char *rfport;        /* "A" (plutogpssim.c:2273) */
    int32_t kernel_buffers;    /* 12 (plutogpssim.c:2102) */
    int32_t pairs_per_push;    /* 300000 */
    const char *iio_lib;       /* NULL: $GPSSINK_IIO_LIB, else libiio.so.0, libiio.so */
    const char *ad9361_lib;    /* NULL: $GPSSINK_AD9361_LIB, else libad9361.so.0, libad9361.so, else the iio library itself */
    int32_t reserved[8];
} gpssink_radio_config;

typedef struct gpssink gpssink;

/* Reference defaults (plutogpssim.c:2270-2276). */
void gpssink_radio_defaults(gpssink_radio_config *cfg);
/* Apply one of the reference's radio options ('A', 'B', 'U', 'N'; 's' sets fs_hz) with its clamping.
 * The strings of -U/-N are NOT copied: they must outlive the open call.  Returns GPSSINK_ERR_ARG for other letters. */
int gpssink_radio_option(gpssink_radio_config *cfg, int letter, const char *arg);

int gpssink_open_null(gpssink **out);
int gpssink_open_file(gpssink **out, const char *path);   /* "-" = stdout */
int gpssink_open_radio(gpssink **out, const gpssink_radio_config *cfg);

/* Synchronous: push `pairs` I/Q pairs.  The radio sink takes whole push units only (pairs % pairs_per_push == 0);
 * each unit is copied into the device buffer and pushed (plutogpssim.c:2148, 2152). */
int gpssink_push(gpssink *s, const int16_t *iq, size_t pairs);

/* Asynchronous: queue `pairs` pairs at `iq` for the sink's writer thread (started on first use) and return a
 * ticket (> 0) or an error (< 0).  `iq` must stay untouched until gpssink_wait(ticket) returns.  Batches are
 * written in submission order.  gpssink_wait returns the status of that batch (an earlier failed batch fails
 * all later ones). */
int64_t gpssink_submit(gpssink *s, const int16_t *iq, size_t pairs);
int gpssink_wait(gpssink *s, int64_t ticket);

/* Stop as fast as the device allows: the writer finishes the push unit it is in (0.1 s of signal on the radio), every
 * batch still queued -- and anything submitted later -- is discarded (its ticket completes with GPSSINK_OK), and
 * gpssink_close then shuts the device down at once.  What the reference does on SIGINT: the TX thread leaves its loop
 * after the current buffer and powers the TX LO down (plutogpssim.c:2014-2022, 2143-2178).  Async-signal-unsafe. */
int gpssink_abort(gpssink *s);

/* Totals so far: pairs accepted by the sink, device pushes (radio) / writes (file). */
int gpssink_stats(gpssink *s, int64_t *pairs, int64_t *pushes);
/* Drains the writer, shuts the radio down the way the reference does (TX LO off, buffer destroyed, channels
 * disabled, context destroyed: plutogpssim.c:2160-2178), closes the file.  Returns the last status. */
int gpssink_close(gpssink *s);
const char *gpssink_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
