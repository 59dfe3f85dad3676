/* TEST INFRASTRUCTURE -- the literal drop-in of INTEGRATION.md, compiled for real.
 *
 * oracle/Makefile target _ref/ref_dropin streams the UNMODIFIED reference source through sed (no copy on disk):
 * the per-sample loop plutogpssim.c:2690-2756 is cut out and replaced by dropin_loop.inc, dropin_init.inc goes in
 * after the I/Q buffer allocation (plutogpssim.c:2604-2609), and this header is force-included.  Everything else --
 * option parsing, RINEX reader, orbit/range/NAV code, the 30 s refresh, the libiio TX thread and its handshake -- is
 * the reference's own code, linked with the capture backend (fake_iio.c) and with libgpsiq.so.  The stream that
 * reaches the backend must be the reference's, byte for byte: tests/test_zz_dropin.py (CPU: against the oracle-backed
 * mock of libgpsiq; -m gpu: against the real CUDA library). */
#include <stdint.h>
#include "gpsiq.h"
static gpsiq_ctx *gq;
static int dropin_prev_prn[GPSIQ_MAX_CHAN];
