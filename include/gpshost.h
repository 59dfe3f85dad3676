/* gpshost.h -- C-ABI of the host orchestrator (libgpshost.so, plain C++ / no CUDA).
 *
 * Produces, epoch by epoch, the channel descriptors (gpsiq_chan_desc, gpsiq.h) that the
 * reference's main loop would hand to its sample loop: it re-implements, from the
 * IS-GPS-200 algorithms and in the reference's evaluation order (so that every double comes
 * out bit-identical under glibc libm), what plutogpssim.c does between two sample loops:
 *
 *   RINEX-2 navigation file (gz)            readRinex2          plutogpssim.c:874-1233
 *   ephemeris set selection / -T overwrite                      plutogpssim.c:2497-2597
 *   satellite position, clock               satpos              plutogpssim.c:443-546
 *   pseudorange, az/el, Klobuchar delay     computeRange        plutogpssim.c:1612-1747
 *   subframes, parity, 6-subframe buffer    eph2sbf, computeChecksum, generateNavMsg
 *                                                               plutogpssim.c:552-814, 1820-1894
 *   channel (re)allocation every 30 s       allocateChannel     plutogpssim.c:1896-1989
 *   per-epoch NCO set-up and gain           computeCodePhase    plutogpssim.c:1754-1787, 2656-2687
 *   user motion file                        readUserMotion      plutogpssim.c:1794-1818
 *
 * SURVEY.md section 8 rows f1 / f3.  The I/Q samples themselves come from libgpsiq.so. */
#ifndef GPSHOST_H
#define GPSHOST_H

#include <stdint.h>

#include "gpsiq.h"

#ifdef __cplusplus
extern "C" {
#endif

#define GPSHOST_OK 0
#define GPSHOST_ERR_ARG (-1)
#define GPSHOST_ERR_NAVFILE (-2)   /* cannot open / not a RINEX navigation file of the selected version */
#define GPSHOST_ERR_NOEPH (-3)     /* no ephemeris (set) usable for the start time */
#define GPSHOST_ERR_MOTION (-4)    /* user motion file missing or empty */
#define GPSHOST_ERR_TIME (-5)      /* start time outside the ephemeris span */

#define GPSHOST_POS_LLH 0   /* -l lat,lon,height [deg,deg,m]   plutogpssim.c:2313-2318 */
#define GPSHOST_POS_XYZ 1   /* -c x,y,z ECEF [m]               plutogpssim.c:2310-2312 */
#define GPSHOST_POS_MOTION 2 /* -u file: t,x,y,z rows at 10 Hz  plutogpssim.c:2301-2304 */

typedef struct gpshost_config {
    const char *nav_path;      /* -e: RINEX-2 (or, with rinex3, RINEX-3) navigation file, plain or gz */
    int32_t pos_mode;          /* GPSHOST_POS_* */
    double pos[3];             /* llh (degrees, metres) or ECEF metres */
    const char *motion_path;   /* GPSHOST_POS_MOTION */
    int32_t have_start;        /* 0: start at the first ephemeris' TOC (reference default) */
    int32_t start[5];          /* -t y,m,d,hh,mm */
    double start_sec;
    int32_t time_overwrite;    /* -T: shift TOC/TOE of the whole file to the start time */
    int32_t iono_disable;      /* -i */
    int64_t sample_rate;       /* -s, Hz (reference default 3000000: TX_SAMPLE_FREQ) */
    int32_t max_chan;          /* 12 = MAX_CHAN, plutogpssim.h:21; up to 32 */
    int32_t carrier_mode;      /* GPSIQ_CARRIER_* */
    int32_t rinex3;            /* -3: nav_path is a RINEX-3 navigation file (plutogpssim.c:1241-1610) */
    int32_t threads;           /* worker threads of gpshost_next over the epochs of a batch: 0 = $GPSHOST_THREADS, else
                                  min(8, cores); 1 = serial.  The descriptors do not depend on it. */
    int32_t reserved[6];
} gpshost_config;

typedef struct gpshost_scenario gpshost_scenario;

int gpshost_open(gpshost_scenario **out, const gpshost_config *cfg);
void gpshost_close(gpshost_scenario *s);
/* Descriptors of the next n_epochs 0.1 s epochs: desc[n_epochs][max_chan].  A slot whose satellite was
 * (re)allocated carries GPSIQ_FLAG_RESET_CARRIER and the initial carrier phase of plutogpssim.c:1964. */
int gpshost_next(gpshost_scenario *s, gpsiq_chan_desc *desc, int n_epochs);
/* Move n_epochs ahead without producing descriptors: gpshost_skip(n) + gpshost_next(m) yields the same descriptors as
 * the last m of gpshost_next(n + m), at the cost of one epoch per 30 s refresh interval instead of every epoch -- how
 * the owner of a later time slice reaches its first epoch.  One difference, deliberate: a
 * channel allocated inside the skipped span does not carry GPSIQ_FLAG_RESET_CARRIER afterwards (its carrier has been
 * running; the phase at a slice boundary is handed over by the previous slice's owner, gpsiq_set_carrier). */
int gpshost_skip(gpshost_scenario *s, int n_epochs);
/* What the reference prints at start-up (plutogpssim.c:2571-2574, 2634-2639): start time and channel table. */
int gpshost_describe(gpshost_scenario *s, char *buf, int buflen);
/* What the reference's -v adds (plutogpssim.c:2487-2495): the ionosphere / UTC parameters of the file header, four
 * lines; empty when the header is incomplete (the reference prints nothing then). */
int gpshost_describe_iono(gpshost_scenario *s, char *buf, int buflen);
/* Receiver time of the next epoch to be produced (GPS week, seconds of week). */
int gpshost_time(gpshost_scenario *s, int *week, double *sec);
const char *gpshost_last_error(void);

/* Known-answer hooks (tests): the same routines the scenario uses. */
uint32_t gpshost_parity(uint32_t source, int nib);                       /* computeChecksum */
void gpshost_date2gps(int y, int m, int d, int hh, int mm, double sec, int *week, double *sow);
void gpshost_llh2xyz(const double llh_rad[3], double xyz[3]);
void gpshost_xyz2llh(const double xyz[3], double llh_rad[3]);

#ifdef __cplusplus
}
#endif
#endif
