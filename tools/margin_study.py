#!/usr/bin/env python3
"""How accurate must a start-phase estimate be, and how accurate is the closed form?  (CPU only)

For one channel and slices of E epochs (the benchmark: 1024 epochs of 300000 samples, groups of 64) the host twin of the
five-level carrier scan (gpsiq_carrier_study_host, the code the device runs) reports
  * the decision margin of the slice-level speculative trajectory: an exact start phase may differ from the estimate by
    at most that much for the slice to be chained by ONE head scan + translation;
  * the smallest group-level margin (what the estimates INSIDE a slice have to meet);
  * the error of the closed-form advance over the slice, raw and with the residual rate the device measures and applies
    (k_bias_update: the residual of the previous slice, per epoch).
usage: margin_study.py [slices] [epochs per slice]"""
import ctypes as C
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pluto_gps_sim_b200 import capi  # noqa: E402


def study(steps, N, T, x0, est_err, rate, gp=64):
    st = np.ascontiguousarray(steps, dtype=np.float64)
    out = np.zeros(5)
    how, fb = C.c_int(0), C.c_int(0)
    capi.check(capi.lib.gpsiq_carrier_study_host(st.ctypes.data, st.size, N, T, float(x0), float(est_err), float(rate), gp,
                                                 out.ctypes.data, C.byref(how), C.byref(fb)))
    return out, how.value, fb.value


def main():
    n_slices = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    E = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    N, T = 300000, 1024
    rng = random.Random(7)
    print("# Doppler Hz | slice margin | min group margin | closed-form error of the slice: raw | with the previous slice's rate | fits from an exact start")
    for fs, f0 in ((2.6e6, 3100.0), (2.6e6, -2400.0), (2.6e6, 420.0), (2.6e6, -60.0), (1e7, 3900.0), (1e7, -1500.0)):
        rate, x0 = 0.0, rng.random()
        for k in range(n_slices):
            # Doppler drifting by ~0.4 Hz/s, as for a satellite well above the horizon
            steps = [(f0 + 0.04 * (k * E + e)) / fs for e in range(E)]
            raw, _, _ = study(steps, N, T, x0, 0.0, 0.0)
            cor, how, fb = study(steps, N, T, x0, 0.0, rate)
            print("%8.1f  %10.2e  %10.2e  %11.2e  %11.2e   how=%d fallbacks=%d unusable groups=%d"
                  % (f0 + 0.04 * k * E, cor[0], cor[3], raw[2], cor[2], how, fb, int(cor[4])))
            rate = raw[2] / E                      # what k_bias_update would measure on this slice (open loop: raw residual per epoch)
            x0 = capi.nco_advance(capi.NCO_CARRIER, x0, steps[0], 0)[0]
            for d in steps:
                x0, _ = capi.nco_advance(capi.NCO_CARRIER, x0, d, N)
        print()


if __name__ == "__main__":
    main()
