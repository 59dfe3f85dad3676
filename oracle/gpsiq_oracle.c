/* TEST INFRASTRUCTURE — CPU restatement of the reference's per-sample
 * synthesis loop (the parity oracle).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this; the
 * product path (pluto_gps_sim_b200/csrc) never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement
 * byte-for-byte against the output of the reference's own source compiled in
 * oracle/_ref (threaded verbatim build and single-thread harness), through the
 * committed goldens in tests/golden/ and, where oracle/_ref is present, live.
 *
 * What is restated (one plain sequential loop, no cleverness on purpose):
 *   sample loop            /root/reference/plutogpssim.c:2689-2756
 *   carrier index          plutogpssim.c:2697 (float) / 2699 (integer #else)
 *   mix + BPSK + gain      plutogpssim.c:2701-2702
 *   accumulate             plutogpssim.c:2705-2706
 *   code NCO + NAV counters plutogpssim.c:2709-2734
 *   chip fetch             plutogpssim.c:2737
 *   carrier NCO            plutogpssim.c:2741-2748
 *   int16 pack             plutogpssim.c:2754-2755
 *   C/A code generator     plutogpssim.c:207-244 (restated as two 10-bit LFSRs)
 *   carrier tables         plutogpssim.c:93-161 (regenerated, see oracle_tables)
 *
 * The input is the per-(epoch, slot) descriptor of include/gpsiq.h — exactly
 * the fields the reference's loop reads, with f_code*delt and f_carr*delt
 * pre-multiplied (the reference recomputes these loop-invariant products on
 * every sample, plutogpssim.c:2709, 2741) and the NAV words reduced to the
 * bit window the epoch can reach.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../include/gpsiq.h"

static int g_sin[512], g_cos[512];
static int g_tables_ready;
static unsigned char g_ca[33][1023];
static int g_ca_ready;

/* The reference's tables (plutogpssim.c:93-161) are reproduced value for value
 * by truncating 511*sin(2*pi*i/512)+1 toward zero -- including the famous
 * asymmetries (sin[0]=1, sin[256]=1, min -510, cos[384]=0 because cos(3pi/2)
 * is a hair below zero in binary64).  tests/test_oracle.py compares these 1024
 * values with the reference's own arrays. */
static void oracle_tables(void) {
    if (g_tables_ready) return;
    const double two_pi = 6.283185307179586476925286766559;
    for (int i = 0; i < 512; i++) {
        g_sin[i] = (int) (511.0 * sin(two_pi * (double) i / 512.0) + 1.0);
        g_cos[i] = (int) (511.0 * cos(two_pi * (double) i / 512.0) + 1.0);
    }
    g_tables_ready = 1;
}

void oracle_get_tables(int32_t *sin_out, int32_t *cos_out) {
    oracle_tables();
    for (int i = 0; i < 512; i++) { sin_out[i] = g_sin[i]; cos_out[i] = g_cos[i]; }
}

/* C/A Gold code, chips as 0/1 (plutogpssim.c:207-244).  G1 = x^10+x^3+1,
 * G2 = x^10+x^9+x^8+x^6+x^3+x^2+1, both seeded all ones; the PRN selects a
 * cyclic delay of G2.  The reference works on +-1 values where -1 is logic 1;
 * its output chip (1 - g1*g2)/2 is the XOR of the two logic levels. */
static void oracle_codegen(unsigned char *ca, int prn) {
    static const int delay[32] = {5, 6, 7, 8, 17, 18, 139, 140, 141, 251, 252, 254, 255, 256, 257, 258,
                                  469, 470, 471, 472, 473, 474, 509, 512, 513, 514, 515, 516, 859, 860, 861, 862};
    unsigned char g1[1023], g2[1023];
    unsigned r1 = 0x3ff, r2 = 0x3ff;          /* bit k = stage k+1 */
    for (int i = 0; i < 1023; i++) {
        g1[i] = (r1 >> 9) & 1;
        g2[i] = (r2 >> 9) & 1;
        unsigned f1 = ((r1 >> 2) ^ (r1 >> 9)) & 1;
        unsigned f2 = ((r2 >> 1) ^ (r2 >> 2) ^ (r2 >> 5) ^ (r2 >> 7) ^ (r2 >> 8) ^ (r2 >> 9)) & 1;
        r1 = ((r1 << 1) | f1) & 0x3ff;
        r2 = ((r2 << 1) | f2) & 0x3ff;
    }
    for (int i = 0, j = 1023 - delay[prn - 1]; i < 1023; i++, j++) ca[i] = g1[i] ^ g2[j % 1023];
}

static void oracle_codes(void) {
    if (g_ca_ready) return;
    for (int p = 1; p <= 32; p++) oracle_codegen(g_ca[p], p);
    g_ca_ready = 1;
}

void oracle_get_ca(int prn, uint8_t *chips1023) {
    oracle_codes();
    memcpy(chips1023, g_ca[prn], 1023);
}

/* One stream of n_epochs epochs.  desc is [n_epochs][nslots]; carr_state
 * (nslots doubles, or uint32 values stored as doubles in integer-carrier mode)
 * is the running carrier phase per slot, read at entry and written at exit --
 * the only quantity the reference carries from one epoch to the next
 * (SURVEY.md §0 item 4).  carr_trace, if not NULL, receives the post-epoch
 * phase for every (epoch, slot).  Returns 0, or -1 on a bad argument. */
int oracle_synth(const gpsiq_chan_desc *desc, int n_epochs, int nslots, int samples_per_epoch,
                 int carrier_mode, double *carr_state, int16_t *iq_out, double *carr_trace) {
    if (!desc || !carr_state || !iq_out || nslots < 1 || nslots > 64) return -1;
    oracle_tables();
    oracle_codes();
    for (int e = 0; e < n_epochs; e++) {
        const gpsiq_chan_desc *d = desc + (size_t) e * nslots;
        int16_t *out = iq_out + (size_t) e * samples_per_epoch * 2;
        /* per-epoch refresh: what computeCodePhase (plutogpssim.c:1754-1787)
         * and the allocation pass (plutogpssim.c:1964) leave in chan[] */
        double code_phase[64], carr_phase[64];
        uint32_t carr_u32[64];
        int icode[64], kbit[64], dataBit[64], codeCA[64];
        for (int c = 0; c < nslots; c++) {
            if (d[c].prn <= 0) continue;
            if (d[c].flags & GPSIQ_FLAG_RESET_CARRIER) carr_state[c] = d[c].carr_phase0;
            carr_phase[c] = carr_state[c];
            carr_u32[c] = (uint32_t) carr_state[c];
            code_phase[c] = d[c].code_phase0;
            icode[c] = d[c].ms0 % 20;
            kbit[c] = 0;
            dataBit[c] = (int) ((d[c].navbits >> kbit[c]) & 1u) * 2 - 1;
            codeCA[c] = (int) g_ca[d[c].prn][(int) code_phase[c]] * 2 - 1;
        }
        for (int n = 0; n < samples_per_epoch; n++) {
            int64_t i_acc = 0, q_acc = 0;
            for (int c = 0; c < nslots; c++) {
                if (d[c].prn <= 0) continue;
                int iTable;
                if (carrier_mode == GPSIQ_CARRIER_FLOAT)
                    iTable = (int) floor(carr_phase[c] * 512.0);
                else
                    iTable = (int) ((carr_u32[c] >> 16) & 0x1ff);
                int ip = dataBit[c] * codeCA[c] * g_cos[iTable] * d[c].gain;
                int qp = dataBit[c] * codeCA[c] * g_sin[iTable] * d[c].gain;
                i_acc += ip;
                q_acc += qp;

                code_phase[c] += d[c].code_step;
                if (code_phase[c] >= 1023.0) {
                    code_phase[c] -= 1023.0;
                    icode[c]++;
                    if (icode[c] >= 20) {
                        icode[c] = 0;
                        kbit[c]++;
                        dataBit[c] = (int) ((d[c].navbits >> kbit[c]) & 1u) * 2 - 1;
                    }
                }
                codeCA[c] = (int) g_ca[d[c].prn][(int) code_phase[c]] * 2 - 1;

                if (carrier_mode == GPSIQ_CARRIER_FLOAT) {
                    carr_phase[c] += d[c].carr_step;
                    if (carr_phase[c] >= 1.0)
                        carr_phase[c] -= 1.0;
                    else if (carr_phase[c] < 0.0)
                        carr_phase[c] += 1.0;
                } else {
                    carr_u32[c] += (uint32_t) (int32_t) d[c].carr_step;
                }
            }
            out[2 * n] = (int16_t) i_acc;
            out[2 * n + 1] = (int16_t) q_acc;
        }
        for (int c = 0; c < nslots; c++) {
            if (d[c].prn <= 0) continue;
            carr_state[c] = (carrier_mode == GPSIQ_CARRIER_FLOAT) ? carr_phase[c] : (double) carr_u32[c];
            if (carr_trace) carr_trace[(size_t) e * nslots + c] = carr_state[c];
        }
    }
    return 0;
}

/* Plain sequential NCO recurrences, used to check the product's exact
 * fast-forward ("jump scan") against the literal per-sample semantics. */
void oracle_code_nco(double phase, double step, int64_t nsteps, double *phase_out, int64_t *wraps_out) {
    int64_t w = 0;
    for (int64_t n = 0; n < nsteps; n++) {
        phase += step;
        if (phase >= 1023.0) { phase -= 1023.0; w++; }
    }
    *phase_out = phase;
    *wraps_out = w;
}

void oracle_carr_nco(double phase, double step, int64_t nsteps, double *phase_out) {
    for (int64_t n = 0; n < nsteps; n++) {
        phase += step;
        if (phase >= 1.0) phase -= 1.0;
        else if (phase < 0.0) phase += 1.0;
    }
    *phase_out = phase;
}
