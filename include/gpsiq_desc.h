/* gpsiq_desc.h -- header-only body of gpsiq_make_desc (include/gpsiq.h), shared by libgpsiq.so and by the
 * host orchestrator (libgpshost.so), which must not depend on the CUDA library to be testable on a CPU box.
 * Folds the fields of the reference's channel_t that its sample loop reads (plutogpssim.c:2697-2746,
 * written at plutogpssim.c:1754-1787, 1964, 2669-2685) into one 64-byte descriptor. */
#ifndef GPSIQ_DESC_H
#define GPSIQ_DESC_H
#include <math.h>
#include <string.h>

#include "gpsiq.h"

static inline int gpsiq_make_desc_inline(gpsiq_chan_desc *out, int carrier_mode, int prn, double f_carr, double f_code,
                                         double delt, double carr_phase, double code_phase, const uint64_t *dwrd60,
                                         int iword, int ibit, int icode, double gain, int carr_phase_is_new) {
    if (!out) return GPSIQ_ERR_ARG;
    memset(out, 0, sizeof *out);
    if (prn <= 0) return GPSIQ_OK;
    if (prn > 32 || !dwrd60 || iword < 0 || ibit < 0 || ibit >= 30 || icode < 0 || icode >= 20) return GPSIQ_ERR_ARG;
    out->prn = prn;
    out->ms0 = iword * 600 + ibit * 20 + icode;
    /* NAV bit window: bit k = data bit number (iword*30 + ibit + k); words past the
       reference's 60-word buffer read as 0 (the reference would over-read, App. A) */
    uint64_t nb = 0;
    /* word by word instead of bit by bit (this loop was a third of the host's per-epoch cost): a NAV word holds its
       30 bits MSB first, the window wants them LSB first -> reverse the word once, drop the bits already sent */
    for (int k = 0, w = iword, skip = ibit; k < 64 && w < 60; w++, skip = 0) {
        uint32_t v = (uint32_t) dwrd60[w] & 0x3FFFFFFFu;
        v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
        v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
        v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
        v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
        v = (v >> 16) | (v << 16);                 /* bit 29 (sent first) is now bit 2 */
        nb |= (uint64_t) (v >> (2 + skip)) << k;   /* bits beyond 64 fall off the top */
        k += 30 - skip;
    }
    out->navbits = nb;
    out->code_phase0 = code_phase;
    volatile double cs = f_code * delt; /* the very product of plutogpssim.c:2709, one rounding */
    volatile double ps = f_carr * delt; /* plutogpssim.c:2741 */
    out->code_step = cs;
    if (carrier_mode == GPSIQ_CARRIER_FLOAT) {
        out->carr_step = ps;
    } else {
        volatile double q = 512.0 * 65536.0 * f_carr;
        q = q * delt; /* left-to-right, plutogpssim.c:2675 */
        out->carr_step = (double) (int) round(q);
    }
    out->carr_phase0 = carr_phase;
    out->gain = gain;
    out->flags = carr_phase_is_new ? GPSIQ_FLAG_RESET_CARRIER : 0u;
    return GPSIQ_OK;
}
#endif
