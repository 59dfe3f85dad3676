"""The safety check behind the line kernel (csrc/line_check.cuh), run on the host:
* minmod == brute-force min of (b + a*x) mod m;
* the literal recurrences of plutogpssim.c:2709-2713 / 2741-2746 never leave the
  error band the check assumes around the straight fixed-point line;
* whenever the line's table/chip index differs from the recurrence's, the check flags the run."""
import random

import numpy as np
import pytest

import oracle_lib as ol
from pluto_gps_sim_b200 import capi


def brute(b, a, m, n):
    x = np.arange(n, dtype=object)
    return int(min((b + a * int(i)) % m for i in range(n)))


def test_minmod_small_exhaustive():
    for m in range(1, 40):
        for a in range(m):
            for b in range(0, m, max(1, m // 7)):
                for n in (1, 2, 3, 5, 11, 64):
                    assert capi.minmod(b, a, m, n) == brute(b, a, m, n), (b, a, m, n)


def test_minmod_random_large_moduli():
    rng = random.Random(5)
    for trial in range(400):
        bits = rng.choice([20, 31, 47, 55, 63])
        m = rng.randrange(2, 1 << bits) if trial % 3 else (1 << bits)
        a = rng.randrange(0, m)
        if trial % 5 == 0:
            a = m - rng.randrange(1, 1000)          # steps just below the modulus
        if trial % 7 == 0:
            a = rng.randrange(0, 1000)              # tiny steps
        b = rng.randrange(0, m)
        n = rng.choice([1, 2, 17, 1024, 5000])
        assert capi.minmod(b, a, m, n) == brute(b, a, m, n), (b, a, m, n)


def test_minmod_early_exit_is_sound():
    rng = random.Random(6)
    for trial in range(300):
        m = 1 << 55
        a, b, n = rng.randrange(m), rng.randrange(m), rng.choice([100, 1024, 4000])
        true_min = brute(b, a, m, n)
        stop = rng.choice([true_min, true_min + 1, true_min // 2, 1 << 40])
        got = capi.minmod(b, a, m, n, stop)
        assert (got < stop) == (true_min < stop), (b, a, n, stop)


FS = [2.6e6, 3.0e6, 10.0e6, 2.1e6]


def test_line_deviation_stays_inside_the_assumed_band():
    rng = random.Random(11)
    n = 1024
    for trial in range(400):
        fs = rng.choice(FS)
        f = rng.uniform(-9000, 9000) if trial % 3 else rng.choice([-1, 1]) * 2.0 ** rng.randint(-20, 13)
        # carrier
        x0 = rng.random() if trial % 6 else rng.choice([0.0, 0.5, 1.0 - 2.0 ** -53, 2.0 ** -40, 0.999999])
        dev, mism, hz = capi.line_probe(capi.NCO_CARRIER, x0, f / fs, n)
        assert dev <= n * 2050 + 2, ("carrier", x0, f / fs, dev)
        assert mism == 0 or hz
        # code
        step = (1.023e6 + f / 1540.0) / fs
        c0 = rng.uniform(0, 1023) if trial % 6 else rng.choice([0.0, 1022.99999999, 1e-9, 511.5, 1022.7])
        dev, mism, hz = capi.line_probe(capi.NCO_CODE, c0, step, n)
        assert dev <= n * 18 + 2, ("code", c0, step, dev)
        assert mism == 0 or hz


def test_near_boundary_lines_are_flagged():
    """Start phases chosen so that the line grazes an index boundary at some sample: the indices may
    or may not differ, but the check must flag every case in which they do -- and flags all grazes."""
    rng = random.Random(12)
    n = 1024
    seen_mismatch = 0
    for trial in range(600):
        fs = rng.choice(FS)
        f = rng.uniform(200, 9000) * rng.choice([-1, 1])
        d = f / fs
        k = rng.randrange(1, n)
        # carrier: phase such that x0 + k*d is within a few 2^-53 of a multiple of 1/512
        target = rng.randrange(0, 512) / 512.0
        x0 = (target - k * d + rng.randrange(-3, 4) * 2.0 ** -53) % 1.0
        dev, mism, hz = capi.line_probe(capi.NCO_CARRIER, x0, d, n)
        assert hz, ("carrier graze not flagged", x0, d, k)
        seen_mismatch += mism
        # code: x0 + k*step within a few 2^-43 of an integer chip (incl. the 1023 wrap)
        step = (1.023e6 + f / 1540.0) / fs
        chip = rng.choice([1023, rng.randrange(1, 1023)])
        c0 = chip - k * step + rng.randrange(-3, 4) * 2.0 ** -43
        if 0.0 <= c0 < 1023.0:
            dev, mism, hz = capi.line_probe(capi.NCO_CODE, c0, step, n)
            assert hz, ("code graze not flagged", c0, step, k)
            seen_mismatch += mism
    assert seen_mismatch > 0   # the probe really produces index disagreements (so `mism == 0 or hz` is not vacuous)


def test_split_word_truncation_is_covered():
    """k_synth_line evaluates the line in split-word form (high word = address bits), which truncates it by
    up to 2^25 (carrier) / 2^19 (code) units: a line that passes just ABOVE a boundary makes the kernel's index
    lag the true one.  Such runs must be flagged, and the lag must really occur in the probe."""
    rng = random.Random(14)
    n = 1024
    lag_seen = 0
    for trial in range(400):
        fs = rng.choice(FS)
        f = rng.uniform(200, 9000) * rng.choice([-1, 1])
        d = f / fs
        k = rng.randrange(17, n)           # past the lane's first sample, where the truncation has built up
        target = rng.randrange(0, 512) / 512.0
        off = rng.randrange(1 << 12, 1 << 21) * 2.0 ** -64   # above the boundary, inside the truncation band
        x0 = (target - k * d + off) % 1.0
        dev, mism, hz = capi.line_probe(capi.NCO_CARRIER, x0, d, n)
        assert mism == 0 or hz, ("carrier", x0, d, k)
        lag_seen += mism
        step = (1.023e6 + f / 1540.0) / fs
        chip = rng.randrange(1, 1023)
        c0 = chip - k * step + rng.randrange(1 << 6, 1 << 15) * 2.0 ** -47
        if 0.0 <= c0 < 1023.0:
            dev, mism, hz = capi.line_probe(capi.NCO_CODE, c0, step, n)
            assert mism == 0 or hz, ("code", c0, step, k)
            lag_seen += mism
    assert lag_seen > 0


def test_generic_lines_are_not_flagged():
    """The check must clear ordinary tiles, or the patch path would carry the load."""
    rng = random.Random(13)
    flagged = 0
    for trial in range(2000):
        d = rng.uniform(-4000, 4000) / 2.6e6
        _, mism, hz = capi.line_probe(capi.NCO_CARRIER, rng.random(), d, 1024)
        flagged += hz
        assert mism == 0 or hz
    assert flagged <= 2


@pytest.mark.parametrize("name,epochs", [("static12", 10), ("circle12", 12), ("allsky32", 4)])
def test_line_kernel_indices_on_reference_descriptors(name, epochs):
    """Whole epochs of reference-derived descriptors (300000 samples each): every sample's table and chip index, as
    k_synth_line evaluates it from its tile anchors (carrier: exact tile-start phase; code: closed form on the epoch's
    line, whose deviation bound grows along the epoch), equals the literal recurrence's in every tile the check
    clears; and the check clears almost all tiles (the patch path must stay a rarity: ~1e-4 of the tiles)."""
    desc = ol.load_golden_desc(name)[:epochs]
    tiles, flagged, bad, _ = capi.line_verify(desc, 300000)
    assert tiles > 0 and bad == 0, (tiles, flagged, bad)
    assert flagged <= max(4, tiles // 4000), (tiles, flagged)
